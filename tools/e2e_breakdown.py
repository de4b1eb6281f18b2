#!/usr/bin/env python3
"""Where the end-to-end overhead around a render goes (host wall clock, one GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from miniraytracer_b200 import api

W, H, SPP = 1920, 1080, int(os.environ.get("SPP", "64"))
hs = api.HostScene(5, W, H)
dev = torch.device("cuda:0")
acc = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
final = torch.empty_like(acc)
host_out = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
stream = torch.cuda.current_stream()
if os.environ.get("RESIDENT", "1") == "1":   # like bench.py: a resident scene that holds its own pool / staging buffers is alive meanwhile
    r0 = api.Renderer(hs, 0); r0.set_stream(stream.cuda_stream); r0.bind_accumulator(acc.data_ptr(), W, H); r0.render_async(W, H, SPP); torch.cuda.synchronize()
for it in range(4):
    torch.cuda.synchronize(); t = [time.perf_counter()]
    r = api.Renderer(hs, 0); torch.cuda.synchronize(); t.append(time.perf_counter())
    r.set_stream(stream.cuda_stream); r.bind_accumulator(acc.data_ptr(), W, H)
    r.render_async(W, H, SPP); t.append(time.perf_counter())
    torch.cuda.synchronize(); t.append(time.perf_counter())
    r.finalize_device(acc.data_ptr(), final.data_ptr(), W, H); torch.cuda.synchronize(); t.append(time.perf_counter())
    host_out.copy_(final, non_blocking=True); torch.cuda.synchronize(); t.append(time.perf_counter())
    st = r.stats(); t.append(time.perf_counter())
    r.close(); t.append(time.perf_counter())
    names = ["create+upload", "launch call", "kernel wait", "finalize", "d2h", "stats", "close"]
    print(it, " ".join(f"{n}={1e3 * (b - a):.2f}ms" for n, a, b in zip(names, t, t[1:])), f"kernel_ms={st['kernel_ms']:.2f}", flush=True)
