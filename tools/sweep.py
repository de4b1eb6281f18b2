#!/usr/bin/env python3
"""Tuning sweep on a GPU box: kernel variants (MRT_MINB launch bounds, MRT_CHUNK pixels per warp task) x workloads.
Prints one JSON line per measurement (kernel time from CUDA events on the launch stream, after a warm-up)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from miniraytracer_b200 import api  # noqa: E402

CASES = {
    "C1": (0, 500, 500, 16), "C1hi": (0, 500, 500, 1024), "C2": (5, 960, 540, 1024), "C2full": (5, 1920, 1080, 1024),
    "C3": (6, 960, 540, 1024), "P_C2": (5, 480, 270, 1024), "P_C1": (0, 256, 256, 1024), "P_C4": (7, 480, 270, 256),
    "P_C5": (8, 480, 270, 256), "F_C2": (5, 1920, 1080, 256), "F_C3": (6, 1920, 1080, 256), "F_C4": (7, 1920, 1080, 64), "F_C5": (8, 3840, 2160, 36),
    "Q_C4": (7, 320, 180, 256), "C4s64": (7, 960, 540, 64), "C4s1k": (7, 960, 540, 1024), "C4s4k": (7, 480, 270, 4096), "C5s1k": (8, 480, 270, 1024), "Q_C5": (8, 320, 180, 256), "Q_C1": (0, 200, 200, 256), "P_C1lo": (0, 500, 500, 16), "C4": (7, 960, 540, 256), "C5": (8, 960, 540, 256),
}


def measure(case, minb, chunk, reps=2, depth=32):
    reps = int(os.environ.get("MRT_SWEEP_REPS", reps))
    scene, w, h, spp = CASES[case]
    os.environ["MRT_MINB"] = str(minb)
    os.environ["MRT_CHUNK"] = str(chunk)
    hs = api.HostScene(scene, w, h)
    r = api.Renderer(hs, 0)
    best = None
    for i in range(reps + 1):
        r.render_async(w, h, spp, depth=depth)
        st = r.stats()
        if i > 0 and (best is None or st["kernel_ms"] < best["kernel_ms"]):
            best = st
    r.close(); hs.close()
    return {"case": case, "minb": minb, "chunk": chunk, "kernel_ms": best["kernel_ms"], "grays_per_s": best["rays"] / best["kernel_ms"] / 1e6,
            "mpaths_per_s": best["paths"] / best["kernel_ms"] / 1e3, "grid": best["grid"], "smem": best["smem_bytes"],
            "alive_frac": best["rays"] / max(1, 32 * best["warp_iterations"])}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="C2")
    ap.add_argument("--minb", default="5")
    ap.add_argument("--chunk", default="0")
    ap.add_argument("--wavefront", default="0")
    ap.add_argument("--all", default="0", help="MRT_VARIANT_ALL values")
    ap.add_argument("--order", default="0")
    ap.add_argument("--binned", default="0", help="MRT_BINNED values (0 off, 1 pool, 2 +box bins, 3 +pending bit)")
    ap.add_argument("--depth", default="32", help="max bounces (comma list)")
    args = ap.parse_args()
    import itertools
    for case, minb, chunk, wf, pf, order in itertools.product(args.cases.split(","), args.minb.split(","), args.chunk.split(","),
                                                             args.wavefront.split(","), args.all.split(","), args.order.split(",")):
      for depth, binned in itertools.product(args.depth.split(","), args.binned.split(",")):
        os.environ["MRT_BINNED"] = binned
        os.environ["MRT_WAVEFRONT"] = wf
        os.environ["MRT_VARIANT_ALL"] = pf
        os.environ["MRT_ORDER"] = order
        res = measure(case, int(minb), int(chunk), depth=int(depth))
        res.update(wavefront=int(wf), variant_all=int(pf), order=int(order), depth=int(depth), binned=int(binned))
        print(json.dumps(res), flush=True)
