#!/usr/bin/env python3
"""How much does the libm canonicalisation matter?  Per BASELINE config (scene at a frame the CPU oracle finishes in seconds, 16 spp,
identical per-(pixel, sample) streams): share of pixels within 1e-4 relative of the GPU render for
  (a) the parity oracle      -- the reference with sinf/cosf/atan2f/asinf/logf/powf bound to correctly rounded versions
                                (oracle/cr_libm.cpp; the GPU uses the same values, csrc/mrt_libm.h), and
  (b) the reference as built -- the same binary with MRT_ORACLE_LIBM=host, i.e. THIS box's glibc float functions.
One JSON line per config.  Needs oracle/_ref/mrt_ref (test infrastructure) and a GPU."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_util
from miniraytracer_b200 import accfile, api

for name, scene, w, h, spp in [("C1", 0, 500, 500, 16), ("C2", 5, 480, 270, 16), ("C3", 6, 480, 270, 16), ("C4", 7, 480, 270, 16), ("C5", 8, 480, 270, 16)]:
    gpu, st = api.render(scene, w, h, spp)
    out = {"config": name, "scene": scene, "frame": [w, h, spp], "gpu_rays": int(st["rays"])}
    for key, host in (("canonical_libm", False), ("host_libm", True)):
        ref, meta = oracle_util.ref_render(scene, w, h, spp, host_libm=host)
        res = accfile.compare(accfile.finalize(gpu), accfile.finalize(ref), rel=1e-4)
        out[key] = {"frac_pixels_within_1e-4": res["frac_ok"], "pixels_off": res["n_bad"], "bit_identical_pixels": float((gpu == ref).all(-1).mean()),
                    "ref_rays": int(meta["rays"]), "rmse": res["rmse"]}
    print(json.dumps(out), flush=True)
