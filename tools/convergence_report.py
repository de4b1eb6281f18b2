#!/usr/bin/env python3
"""RMSE of a GPU render against the oracle vs the oracle's own seed-to-seed noise floor (north_star check B), one JSON
line per case.  Needs oracle/_ref/mrt_ref (test infrastructure) and a GPU."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle_util
from miniraytracer_b200 import accfile, api

def rmse(a, b):
    return float(np.sqrt(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2)))

for scene, w, h, spp in [(5, 96, 54, 4096), (7, 64, 36, 4096), (8, 64, 36, 4096), (6, 128, 72, 1024), (0, 100, 100, 256)]:
    sa, sb = oracle_util.DEFAULT_SEED, 987654321
    ref_a = accfile.finalize(oracle_util.ref_render(scene, w, h, spp, seed=sa)[0])
    ref_b = accfile.finalize(oracle_util.ref_render(scene, w, h, spp, seed=sb)[0])
    hs = api.HostScene(scene, w, h); r = api.Renderer(hs, 0)
    r.render_async(w, h, spp, seed=sa); gpu_a = accfile.finalize(r.readback())
    r.close(); hs.close()
    print(json.dumps({"scene": scene, "frame": [w, h, spp], "noise_floor_rmse(ref_a,ref_b)": rmse(ref_a, ref_b),
                      "rmse(gpu_a,ref_b)": rmse(gpu_a, ref_b), "rmse(gpu_a,ref_a)": rmse(gpu_a, ref_a)}), flush=True)
