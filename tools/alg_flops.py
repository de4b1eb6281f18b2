#!/usr/bin/env python3
"""Algorithmic flop per ray of the five BASELINE configs, re-derived from operation counters.

The roofline's "achieved" figure counts the REFERENCE ALGORITHM's arithmetic: sum over the operations the reference performs
(op count x flops per op, the per-op table below is SURVEY 8d's), independent of how the GPU code is written.  The counts come
from the test-side host build of the device tracer core (tests/host_emul/emul_render -counters: trace_core.h compiled by g++ with
MRT_COUNT_OPS; its ray / box / primitive counts equal the reference's -- same trace() count as the oracle on every scene), run
WITHOUT the flattener's translate cull boxes (-nocull), i.e. counting every transform and rotated-box test the reference makes.

    python tools/alg_flops.py            # prints one JSON line per config + a table; --write updates tools/alg_flops.json

bench.py reads tools/alg_flops.json (committed) for its roofline.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# flops per operation (1 flop per add/sub/mul/div/sqrt/min/max/cmp, 1 per transcendental, RNG step = 12)
WEIGHTS = {
    "ray_ctor": 9, "aabb": 22, "sphere": 18, "sphere_hit": 22, "sphere_moving": 12, "rect": 18, "rect_hit": 12, "tri": 44, "tri_hit": 44,
    "vol": 6, "translate": 4, "rotate": 16, "rng": 12, "lambert": 81, "metal": 45, "dielectric": 55, "isotropic": 30, "lightpdf": 21,
    "perlin": 213, "image": 18, "checker": 9, "sky": 11, "paths": 56,
}
# config -> (scene, width, height, spp) of the counting run (op counts per ray do not depend on the frame size beyond noise)
RUNS = {"C1": (0, 250, 250, 16), "C2": (5, 320, 180, 16), "C3": (6, 320, 180, 16), "C4": (7, 320, 180, 16), "C5": (8, 320, 180, 16)}
SURVEY = {"C1": 832.0, "C2": 294.0, "C3": 336.0, "C4": 813.0, "C5": 1258.0}   # the survey's probe figures, for comparison


def derive(name):
    import oracle_util
    exe = oracle_util.build_emul()
    scene, w, h, spp = RUNS[name]
    _, meta = oracle_util.emul_render(exe, scene, w, h, spp, extra=["-nocull"])
    c = meta["counters"]
    flops = sum(WEIGHTS[k] * c[k] for k in WEIGHTS)
    rays = c["rays"]
    return {"config": name, "scene": scene, "frame": [w, h, spp], "rays": rays, "rays_per_path": rays / c["paths"],
            "flop_per_ray": flops / rays, "flop_per_path": flops / c["paths"], "survey_flop_per_ray": SURVEY[name],
            "ratio_to_survey": flops / rays / SURVEY[name],
            "per_ray": {k: c[k] / rays for k in WEIGHTS}}


if __name__ == "__main__":
    out = {}
    for name in RUNS:
        r = derive(name)
        out[name] = r
        print(json.dumps(r))
    print("\nconfig  flop/ray  survey  ratio  rays/path  aabb/ray  sphere/ray  rect/ray  tri/ray  rng/ray")
    for n, r in out.items():
        p = r["per_ray"]
        print(f"{n:6s} {r['flop_per_ray']:9.1f} {r['survey_flop_per_ray']:7.0f} {r['ratio_to_survey']:6.3f} {r['rays_per_path']:10.2f} "
              f"{p['aabb']:9.2f} {p['sphere']:11.2f} {p['rect']:9.2f} {p['tri']:8.2f} {p['rng']:8.2f}")
    if "--write" in sys.argv:
        json.dump({"weights": WEIGHTS, "configs": out}, open(os.path.join(ROOT, "tools", "alg_flops.json"), "w"), indent=1)
