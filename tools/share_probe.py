#!/usr/bin/env python3
"""Kernel time of one sample slice of a frame: scene:W:H:spp:s_begin:s_end[,...] (the per-GPU share of a multi-GPU render, run alone).
Separates a fixed cost per launch from a per-path cost that depends on the samples per pixel of the slice."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from miniraytracer_b200 import api  # noqa: E402

for case in sys.argv[1].split(","):
    scene, w, h, spp, b, e = (int(v) for v in case.split(":"))
    tuning = dict(kv.split("=") for kv in sys.argv[2:])
    hs = api.HostScene(scene, w, h)
    r = api.Renderer(hs, 0, {k: int(v) for k, v in tuning.items()})
    best = None
    for i in range(3):
        r.render_async(w, h, spp, depth=32, sample_begin=b, sample_end=e)
        st = r.stats()
        if i and (best is None or st["kernel_ms"] < best["kernel_ms"]):
            best = st
    r.close(); hs.close()
    paths = w * h * (e - b)
    print(json.dumps({"case": case, "kernel_ms": round(best["kernel_ms"], 3), "ns_per_path": round(1e6 * best["kernel_ms"] / paths, 5),
                      "alive_frac": round(best["rays"] / max(1, 32 * best["warp_iterations"]), 4),
                      "busy": round(best["warp_time_sum_ns"] / max(1, best["warps"] * best["warp_span_ns"]), 4),
                      "stage_sum_frac": round(best["stage_sum_ns"] / max(1, best["warp_time_sum_ns"]), 4), "tuning": tuning}), flush=True)
