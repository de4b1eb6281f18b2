#!/usr/bin/env python3
"""Golden crop windows at BASELINE.json's TRUE configuration sizes, generated from the reference oracle
(oracle/_ref/mrt_ref render ... -x0 -x1 -y0 -y1: the reference's own get_ray + trace per (pixel, sample), only for the
pixels of the window; PCG32 stream ids and sub-pixel positions are those of the full frame).  Run in the build container:

    python tests/golden/make_golden_windows.py

window_<name>.npz : acc[h,w,4] float32 of the window + the render parameters; used by
tests/test_gpu_parity.py::test_true_size_window_parity.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_util  # noqa: E402
from test_gpu_parity import WINDOWS  # noqa: E402

if __name__ == "__main__":
    assert oracle_util.ensure_ref(), "oracle/_ref/mrt_ref not available"
    for name, scene, W, H, spp, crop in WINDOWS:
        acc, meta = oracle_util.ref_render(scene, W, H, spp, crop=crop)
        x0, y0, x1, y1 = crop
        max_stream = ((y1 - 1) * W + (x1 - 1)) * spp + spp - 1
        np.savez_compressed(os.path.join(HERE, f"window_{name}.npz"), acc=acc, scene=scene, width=W, height=H, spp=spp, depth=32,
                            crop=np.array(crop, dtype=np.uint32), seed=np.uint64(oracle_util.DEFAULT_SEED), rays=np.uint64(meta["rays"]),
                            max_stream=np.uint64(max_stream))
        print(name, "rays", meta["rays"], "max stream id 2^%.2f" % np.log2(max_stream), "mean", acc[..., :3].sum() / max(1.0, acc[..., 3].sum()))
