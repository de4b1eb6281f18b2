#!/usr/bin/env python3
"""Tuning sweep on a GPU box: MrtTuning fields (launch bounds, chunk size, mode, bins, ...) x workloads.
Prints one JSON line per measurement (kernel time from CUDA events on the launch stream, after a warm-up)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from miniraytracer_b200 import api  # noqa: E402

CASES = {
    "C2n8": (5, 1920, 1080, 128), "C3n8": (6, 1920, 1080, 128), "C4n8": (7, 1920, 1080, 512), "C5n8": (8, 1920, 1080, 256),   # the per-GPU share of an 8-GPU run
    "N_C4": (7, 1920, 1080, 64), "N_C5": (8, 3840, 2160, 16), "N_C3": (6, 1920, 1080, 64),   # bench resolution, few samples: ncu captures
    "C1": (0, 500, 500, 16), "C1hi": (0, 500, 500, 1024), "C2": (5, 960, 540, 1024), "C2full": (5, 1920, 1080, 1024),
    "C3": (6, 960, 540, 1024), "P_C2": (5, 480, 270, 1024), "P_C1": (0, 256, 256, 1024), "P_C4": (7, 480, 270, 256),
    "P_C5": (8, 480, 270, 256), "F_C2": (5, 1920, 1080, 256), "F_C3": (6, 1920, 1080, 256), "F_C4": (7, 1920, 1080, 64), "F_C5": (8, 3840, 2160, 36),
    "Q_C4": (7, 320, 180, 256), "C4s64": (7, 960, 540, 64), "C4s1k": (7, 960, 540, 1024), "C4s4k": (7, 480, 270, 4096), "C5s1k": (8, 480, 270, 1024), "Q_C5": (8, 320, 180, 256), "Q_C1": (0, 200, 200, 256), "P_C1lo": (0, 500, 500, 16), "C4": (7, 960, 540, 256), "C5": (8, 960, 540, 256),
}


def measure(case, tuning, reps=2, depth=32):
    reps = int(os.environ.get("MRT_SWEEP_REPS", reps))
    scene, w, h, spp = CASES[case] if case in CASES else tuple(int(v) for v in case.split(":"))   # ad hoc: scene:W:H:spp
    hs = api.HostScene(scene, w, h)
    r = api.Renderer(hs, 0, tuning)
    best = None
    for i in range(reps + 1):
        r.render_async(w, h, spp, depth=depth)
        st = r.stats()
        if i > 0 and (best is None or st["kernel_ms"] < best["kernel_ms"]):
            best = st
    r.close(); hs.close()
    res = {"case": case, "kernel_ms": best["kernel_ms"], "grays_per_s": best["rays"] / best["kernel_ms"] / 1e6,
           "mpaths_per_s": best["paths"] / best["kernel_ms"] / 1e3, "grid": best["grid"], "smem": best["smem_bytes"], "ran_mode": best["mode"], "ran_coop": best["coop_trees"],
           "alive_frac": best["rays"] / max(1, 32 * best["warp_iterations"]), "depth": depth,
           "coop_node_fill": best["coop_node_items"] / max(1, 32 * best["coop_node_steps"]),
           "coop_leaf_fill": best["coop_leaf_items"] / max(1, 32 * best["coop_leaf_steps"]),
           "coop_node_items_per_ray": best["coop_node_items"] / max(1, best["rays"])}
    if best.get("warps"):
        res["warp_busy_frac"] = best["warp_time_sum_ns"] / max(1, best["warps"] * best["warp_span_ns"])
        res["warp_span_ms"] = best["warp_span_ns"] / 1e6
        res["first_exit_ms"] = best["first_exit_ns"] / 1e6
    res["tuning"] = {k: v for k, v in tuning.items() if v}
    return res


if __name__ == "__main__":
    import itertools
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="C2")
    ap.add_argument("--depth", default="32", help="max bounces (comma list)")
    fields = [n for n, _ in api.Tuning._fields_ if n != "reserved"]
    for f in fields:   # every MrtTuning field as a comma list, e.g. --min_blocks 5,6 --coop_trees 1,2
        ap.add_argument("--" + f, default="0")
    args = ap.parse_args()
    for case, depth in itertools.product(args.cases.split(","), args.depth.split(",")):
        for combo in itertools.product(*[getattr(args, f).split(",") for f in fields]):
            tuning = {f: int(v) for f, v in zip(fields, combo)}
            print(json.dumps(measure(case, tuning, depth=int(depth))), flush=True)
