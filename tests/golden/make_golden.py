#!/usr/bin/env python3
"""Generates the committed golden fixtures from the reference oracle (oracle/_ref/mrt_ref, i.e. the
reference's own code with per-(pixel, sample) RNG streams).  Run in the build container:

    python tests/golden/make_golden.py

golden_sceneS.npz : acc[h,w,4] float32 = (sum of finite samples, count), plus the render parameters.
kat.txt           : PCG32 / sampler / Perlin known-answer vectors printed by `mrt_ref kat`.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_util  # noqa: E402

W, H, SPP, DEPTH = 64, 36, 4, 32

if __name__ == "__main__":
    assert oracle_util.ensure_ref(), "oracle/_ref/mrt_ref not available"
    for scene in range(9):
        acc, meta = oracle_util.ref_render(scene, W, H, SPP, DEPTH)
        np.savez_compressed(os.path.join(HERE, f"golden_scene{scene}.npz"), acc=acc, scene=scene, width=W, height=H,
                            spp=SPP, depth=DEPTH, seed=np.uint64(oracle_util.DEFAULT_SEED), rays=np.uint64(meta["rays"]))
        print("scene", scene, "rays", meta["rays"])
    # light list with both allocated entries (`-lights all`): the sphere-light importance sampling path (sphere.cpp:63-79)
    for scene in (5, 7):
        acc, meta = oracle_util.ref_render(scene, W, H, 16, DEPTH, all_lights=True)
        np.savez_compressed(os.path.join(HERE, f"golden_all_lights_scene{scene}.npz"), acc=acc, scene=scene, width=W, height=H,
                            spp=16, depth=DEPTH, seed=np.uint64(oracle_util.DEFAULT_SEED), rays=np.uint64(meta["rays"]))
    # Cornell box + two triangle_scene_objects (`-extra triangles`): the lone-triangle class (triangle.cpp:5-175)
    acc, meta = oracle_util.ref_render(5, W, H, 16, DEPTH, extra_triangles=True)
    np.savez_compressed(os.path.join(HERE, "golden_extra_triangles_scene5.npz"), acc=acc, scene=5, width=W, height=H, spp=16, depth=DEPTH,
                        seed=np.uint64(oracle_util.DEFAULT_SEED), rays=np.uint64(meta["rays"]))
    kat = oracle_util.ref_run(["kat"]).stdout
    open(os.path.join(HERE, "kat.txt"), "w").write(kat)
    # scene dumps printed by the reference + matching renders, used to test oracle/mrt_oracle.c without the reference binary
    for scene in (2, 3, 5, 6):
        oracle_util.ref_dump_scene(scene, 200, 160, os.path.join(HERE, f"scene{scene}_dump_200x160.txt"))
        acc, meta = oracle_util.ref_render(scene, 200, 160, 4)
        np.savez_compressed(os.path.join(HERE, f"golden_restatement_scene{scene}.npz"), acc=acc, rays=np.uint64(meta["rays"]))
