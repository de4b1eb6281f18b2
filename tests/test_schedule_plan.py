"""The mode-B ticket-queue planner (miniraytracer_b200/csrc/schedule.h, called by mrt_gpu_render_async) on the CPU:
tests/host_emul/check_plan.cpp walks every ticket of a plan with the kernel's task -> pixels mapping; the tickets must tile the
frame exactly, in order, every chunk within the staging capacity.  Host logic only -- no GPU, no oracle."""
import json
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAX_ITEMS = 8192   # kMaxStageItems (render_kernels.cuh)


@pytest.fixture(scope="module")
def check_plan(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("plan") / "check_plan")
    subprocess.run(["g++", "-std=c++20", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "miniraytracer_b200", "csrc"),
                    "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "host_emul", "check_plan.cpp"), "-o", exe], check=True)
    return exe


def _run(exe, cases):
    text = "".join(" ".join(str(int(v)) for v in c) + "\n" for c in cases)
    r = subprocess.run([exe], input=text, capture_output=True, text=True)
    out = [json.loads(l) for l in r.stdout.strip().splitlines()]
    assert r.returncode == 0 and len(out) == len(cases), out[-1:]
    return out


def test_plans_of_the_baseline_configs(check_plan):
    """(n_pixels, ns, resident warps, trees, cooperative): the five BASELINE configs on 1 and 8 GPUs."""
    cases = []
    for ns_div in (1, 8):
        cases += [(500 * 500, 16 // ns_div, 2960, 1, 0, 0, 0, 0, MAX_ITEMS), (1920 * 1080, 1024 // ns_div, 3552, 0, 0, 0, 0, 0, MAX_ITEMS),
                  (1920 * 1080, 4096 // ns_div, 2960, 1, 0, 0, 0, 0, MAX_ITEMS), (3840 * 2160, 4096 // ns_div, 2960, 1, 1, 0, 0, 0, MAX_ITEMS)]
    out = _run(check_plan, cases)
    for c, o in zip(cases, out):
        n_pixels, ns, warps = c[0], c[1], c[2]
        assert o["biggest"] * ns <= MAX_ITEMS
        # several tickets per resident warp, unless the chunks are already at their minimum of 256 paths (tiny launches)
        assert o["n_tasks"] >= min(n_pixels, 10 * warps) or o["K"] * ns <= max(256, ns)
        assert o["last"] * ns <= max(ns, o["K"] * ns // 16 + ns)   # the launch ends on small chunks
    c2_full, c2_slice = out[1], out[5]
    assert c2_full["K"] * 1024 == 8192 and c2_slice["K"] * 128 <= 8192   # list scenes: chunks up to the staging limit
    c4_slice = out[6]
    assert c4_slice["K"] * 512 == 2048                                     # per-lane tree scenes: 2048-path chunks


def test_random_and_edge_plans_tile_the_frame(check_plan):
    rnd = random.Random(20261019)
    cases = [(1, 1, 3552, 0, 0, 0, 0, 0, MAX_ITEMS), (7, 8192, 3552, 1, 1, 0, 0, 0, MAX_ITEMS), (31, 1, 4, 0, 0, 0, 0, 0, MAX_ITEMS),
             (100, 3, 4, 0, 0, 5, 0, 0, MAX_ITEMS), (2 ** 31 - 1, 1, 3552, 0, 0, 0, 0, 0, MAX_ITEMS)][:4]
    for _ in range(400):
        n_pixels = rnd.choice([rnd.randint(1, 64), rnd.randint(65, 5000), rnd.randint(5001, 300000)])
        ns = rnd.choice([1, 2, 3, 16, 31, 32, 33, 100, 121, 128, 129, 256, 1000, 1024, 4096, 8192])
        warps = rnd.choice([1, 4, 148, 2960, 3552, 4736])
        trees, coop = rnd.choice([(0, 0), (1, 0), (1, 1)])
        chunk_pixels = rnd.choice([0, 0, 0, 1, 3, 64])
        chunk_paths = rnd.choice([0, 0, 256, 512, 4096, 8192, 100000])
        tail_tasks = rnd.choice([0, 0, 1, 8])
        cases.append((n_pixels, ns, warps, trees, coop, chunk_pixels, chunk_paths, tail_tasks, MAX_ITEMS))
    _run(check_plan, cases)
