#!/usr/bin/env python3
"""torchrun check of the progressive multi-GPU path (NCCL): every rank renders its passes into a bound accumulator,
ProgressiveReducer overlaps the per-pass all-reduce with the next pass; rank 0 compares the final reduced accumulator
with a one-shot single-GPU render of all samples and prints one JSON line.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/progressive_check.py
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from miniraytracer_b200 import api, accfile, distributed as mdist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, SPP, PASSES = 960, 540, 1024, 8
hs = api.HostScene(5, W, H)
r = api.Renderer(hs, local)
acc = torch.zeros((H, W, 4), dtype=torch.float32, device=f"cuda:{local}")
r.set_stream(torch.cuda.current_stream().cuda_stream)
r.bind_accumulator(acc.data_ptr(), W, H)
t_prev = []
t0 = time.perf_counter()
red = mdist.ProgressiveReducer(acc, lambda b, e, out: r.render_async(W, H, SPP, sample_begin=b, sample_end=e, accumulate=True),
                               preview=lambda p, buf: t_prev.append(time.perf_counter() - t0))
final = red.run(mdist.progressive_schedule(SPP, rank, world, PASSES))
torch.cuda.synchronize()
t_total = time.perf_counter() - t0
got = final.cpu().numpy()
if rank == 0:
    r2 = api.Renderer(hs, local)
    r2.render_async(W, H, SPP)
    full = r2.readback()
    r2.close()
    res = accfile.compare(accfile.finalize(got), accfile.finalize(full), rel=1e-4)
    print(json.dumps({"world": world, "passes": PASSES, "frame": [W, H, SPP], "total_s": t_total, "preview_times_s": t_prev,
                      "counts_equal": bool((got[..., 3] == full[..., 3]).all()), "frac_ok": res["frac_ok"], "n_bad": int(res["n_bad"])}), flush=True)
r.close(); hs.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
