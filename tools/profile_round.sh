#!/bin/bash
# One GPU-box profiling pass (run under gpurun): GPU tests, the default bench line, the ncu launch list of the bench command and
# ncu --set full captures of the render kernel on every BASELINE config at bench resolution.  Outputs land in gpurun_out/.
export MRT_NO_BUILD=1 MRT_SWEEP_REPS=1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err; tail -1 gpurun_out/r2p_bench_n1.err
B="python bench.py --no-per-config --no-cpu-baseline --steps 2 --warmup 1"
$B > gpurun_out/plain_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2p_ncu_launch_list_bench.csv $B > gpurun_out/ncu_ll.log 2>&1
$B > gpurun_out/plain_b2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_pixel -s 2 -c 1 -o gpurun_out/r2p_full_C2_bench $B > gpurun_out/ncu_full.log 2>&1
for c in C1 N_C3 N_C4 N_C5; do
  S="python tools/sweep.py --cases $c"
  $S > gpurun_out/plain_$c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_pixel -s 1 -c 1 -o gpurun_out/r2p_full_$c $S > gpurun_out/ncu_$c.log 2>&1
done
S="python tools/sweep.py --cases N_C4 --coop_trees 2"; $S > gpurun_out/plain_c4c.log 2>&1 && ncu --set full --clock-control none -k regex:render_pixel -s 1 -c 1 -o gpurun_out/r2p_full_N_C4_coop $S > gpurun_out/ncu_c4c.log 2>&1
S="python tools/sweep.py --cases C1 --coop_trees 2"; $S > gpurun_out/plain_c1c.log 2>&1 && ncu --set full --clock-control none -k regex:render_pixel -s 1 -c 1 -o gpurun_out/r2p_full_C1_coop $S > gpurun_out/ncu_c1c.log 2>&1
S="python tools/sweep.py --cases N_C5 --coop_trees 1"; $S > gpurun_out/plain_c5l.log 2>&1 && ncu --set full --clock-control none -k regex:render_pixel -s 1 -c 1 -o gpurun_out/r2p_full_N_C5_lane $S > gpurun_out/ncu_c5l.log 2>&1
# the SEQ instantiation on the share of one GPU of an 8-GPU render (C2, 128 of 1024 samples)
S="python tools/share_probe.py 5:1920:1080:1024:0:128"; $S > gpurun_out/plain_seq.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:render_pixel -s 1 -c 1 -o gpurun_out/r2p_full_C2_slice128 $S > gpurun_out/ncu_seq.log 2>&1
cat gpurun_out/plain_C1.log gpurun_out/plain_N_C3.log gpurun_out/plain_N_C4.log gpurun_out/plain_N_C5.log gpurun_out/plain_c4c.log gpurun_out/plain_c1c.log gpurun_out/plain_c5l.log | cut -c1-110
ls -la gpurun_out/*.ncu-rep | tail -12
