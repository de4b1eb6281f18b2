#!/usr/bin/env python3
"""Golden pairs for the tone-map parity test: the reference's OWN main() (`mrt_ref stock`, its own threading and RNG) renders a
small frame, and the harness dumps G_linearBackBuffer (the final linear image) together with G_backBuffer (the reference's adaptive
logarithmic tone map of exactly that image, main.cpp:416-444, packed by ARGB32, vec3.h:327-333).  Run in the build container:

    python tests/golden/make_golden_tonemap.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_util  # noqa: E402
from miniraytracer_b200.accfile import read_acc  # noqa: E402

if __name__ == "__main__":
    assert oracle_util.ensure_ref()
    for scene, w, h, spp in [(5, 96, 54, 64), (0, 80, 80, 16), (7, 96, 54, 16)]:
        with tempfile.TemporaryDirectory() as d:
            lin, argb = os.path.join(d, "lin.bin"), os.path.join(d, "argb.u32")
            oracle_util.ref_run(["stock", "-scene", scene, "-width", w, "-height", h, "-samples", spp, "-depth", 32, "-mode", 0, "-threads", 4,
                                 "-dump", lin, "-dumpargb", argb])
            acc, _ = read_acc(lin)
            img = np.fromfile(argb, dtype=np.uint32).reshape(h, w)
        np.savez_compressed(os.path.join(HERE, f"tonemap_scene{scene}.npz"), linear=acc[..., :3].copy(), argb=img, width=w, height=h)
        print("scene", scene, "max luminance", float((acc[..., :3] * np.array([0.212655, 0.715158, 0.072187], np.float32)).sum(-1).max()),
              "argb mean", float(((img >> 8) & 255).mean()))
