#!/bin/bash
# One GPU-box profiling pass (run under gpurun): the ncu launch list of the bench command and ncu --set full captures of the
# render kernel (the bench command itself = the lane-strided instantiation; the 128-sample slice of C2 = the SEQ instantiation,
# i.e. the share of one GPU of an 8-GPU render).  Reports are converted to CSV here and deleted: gpurun_out/ carries <= 64 MiB back.
# Usage: tools/profile_round.sh [prefix]   (per-config captures at bench resolution: tools/sweep.py --cases C1,N_C3,N_C4,N_C5 under
# the same ncu command line; the round-2 set is profiles/r2i_ncu_full_*.csv)
P=${1:-r2p}
export MRT_NO_BUILD=1 MRT_SWEEP_REPS=1
B="python bench.py --no-per-config --no-cpu-baseline --steps 2 --warmup 1"
$B > gpurun_out/plain_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${P}_ncu_launch_list_bench.csv $B > gpurun_out/ncu_ll.log 2>&1
$B > gpurun_out/plain_b2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_pixel -s 2 -c 1 -o gpurun_out/${P}_full_C2_bench $B > gpurun_out/ncu_full.log 2>&1
S="python tools/share_probe.py 5:1920:1080:1024:0:128"
$S > gpurun_out/plain_seq.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_pixel -s 1 -c 1 -o gpurun_out/${P}_full_C2_slice128 $S > gpurun_out/ncu_seq.log 2>&1
for r in gpurun_out/${P}_full_*.ncu-rep; do
  ncu -i $r --page raw --csv > ${r%.ncu-rep}_raw.csv 2>/dev/null
  rm -f $r
done
ls -la gpurun_out/ | tail -12
