"""Test-side helpers around the oracle (oracle/_ref/mrt_ref = the patched reference renderer).

Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may use anything under oracle/.
"""
import hashlib
import json
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "mrt_ref")
ASSETS = os.path.join(ROOT, "assets")
RUN_DIR = os.path.join(ASSETS, "run")
CACHE = os.path.join(tempfile.gettempdir(), "mrt_oracle_cache")
CSRC = os.path.join(ROOT, "miniraytracer_b200", "csrc")
DEFAULT_SEED = 11350390909718046443


def have_ref():
    return os.path.exists(REF_BIN) and os.path.isdir(RUN_DIR)


def ensure_ref():
    """Build oracle/_ref/mrt_ref if the reference sources are available (this container only)."""
    if not have_ref() and os.path.isdir(os.environ.get("MRT_REFERENCE_DIR", "/root/reference")):
        subprocess.run([os.path.join(ROOT, "oracle", "build_ref.sh")], check=True)
    return have_ref()


def ref_run(args, **kw):
    return subprocess.run([REF_BIN] + [str(a) for a in args], cwd=RUN_DIR, check=True, capture_output=True, text=True, **kw)


def ref_render(scene, width, height, spp, depth=32, seed=DEFAULT_SEED, s0=0, s1=0, threads=0, crop=None, all_lights=False, host_libm=False, draw2=False, maxlum=None, extra_triangles=False):
    """Oracle render with per-(pixel, sample) RNG streams; returns (acc[h,w,4], meta). Cached in /tmp.
    crop = (x0, y0, x1, y1): only that window of the frame (stream ids and u,v stay the full frame's)."""
    from miniraytracer_b200.accfile import read_acc
    os.makedirs(CACHE, exist_ok=True)
    st = os.stat(REF_BIN)
    key = hashlib.sha1(f"{scene}-{width}-{height}-{spp}-{depth}-{seed}-{s0}-{s1}-{crop}-{all_lights}-{host_libm}-{draw2}-{maxlum}-{extra_triangles}-{st.st_size}-{int(st.st_mtime)}".encode()).hexdigest()[:16]
    path = os.path.join(CACHE, f"ref_{key}.bin")
    if not os.path.exists(path):
        tmp = path + f".{os.getpid()}.tmp"
        ref_run(["render", "-scene", scene, "-width", width, "-height", height, "-samples", spp, "-depth", depth,
                 "-seed", seed, "-s0", s0, "-s1", s1, "-threads", threads, "-out", tmp] +
                (["-x0", crop[0], "-y0", crop[1], "-x1", crop[2], "-y1", crop[3]] if crop else []) + (["-lights", "all"] if all_lights else []) +
                (["-draw2", 1] if draw2 else []) + (["-extra", "triangles"] if extra_triangles else []) + (["-maxlum", maxlum] if maxlum is not None else []),
                **({"env": dict(os.environ, MRT_ORACLE_LIBM="host")} if host_libm else {}))   # host: the box's own libm instead of the canonical one
        os.replace(tmp, path)
    return read_acc(path)


def ref_dump_scene(scene, width, height, out, all_lights=False, extra_triangles=False):
    ref_run(["dump-scene", "-scene", scene, "-width", width, "-height", height, "-out", out] + (["-lights", "all"] if all_lights else []) +
            (["-extra", "triangles"] if extra_triangles else []))


def build_emul():
    """g++ build of the TEST-ONLY host emulation of the device tracer core."""
    out_dir = os.path.join(ROOT, "build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "emul_render")
    srcs = [os.path.join(ROOT, "tests", "host_emul", "emul_render.cpp")] + \
           [os.path.join(CSRC, f) for f in ("scene_graph.cpp", "scenes.cpp", "obj_loader.cpp", "flatten.cpp")]
    deps = srcs + [os.path.join(CSRC, "trace_core.h"), os.path.join(CSRC, "scene_graph.h"), os.path.join(ROOT, "include", "mrt_types.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.run(["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-I", CSRC, "-I", os.path.join(ROOT, "include")] + srcs +
                       ["-o", exe, "-lpthread"], check=True)
    return exe


def emul_render(exe, scene, width, height, spp, depth=32, seed=DEFAULT_SEED, s0=0, s1=0, extra=(), crop=None):
    from miniraytracer_b200.accfile import read_acc
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        path = f.name
    try:
        r = subprocess.run([exe, "-scene", str(scene), "-width", str(width), "-height", str(height), "-samples", str(spp),
                            "-depth", str(depth), "-seed", str(seed), "-s0", str(s0), "-s1", str(s1), "-assets", ASSETS,
                            "-out", path, "-counters"] + list(extra) +
                           (["-x0", str(crop[0]), "-y0", str(crop[1]), "-x1", str(crop[2]), "-y1", str(crop[3])] if crop else []),
                           check=True, capture_output=True, text=True)
        acc, meta = read_acc(path)
        meta["counters"] = json.loads(r.stdout.strip().splitlines()[-1])
        return acc, meta
    finally:
        os.unlink(path)


def build_emul_binned():
    """g++ build of tests/host_emul/emul_binned.cpp: the render kernels of render_kernels.cuh compiled for the CPU through
    the SIMT shim emul_warp.h (lanes = fibers).  TEST-ONLY; needs the CUDA headers for the vector types."""
    out_dir = os.path.join(ROOT, "build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "emul_binned")
    srcs = [os.path.join(ROOT, "tests", "host_emul", "emul_binned.cpp")] + \
           [os.path.join(CSRC, f) for f in ("scene_graph.cpp", "scenes.cpp", "obj_loader.cpp", "flatten.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("trace_core.h", "render_kernels.cuh", "gpu_internal.h", "scene_graph.h")] + \
           [os.path.join(ROOT, "tests", "host_emul", "emul_warp.h"), os.path.join(ROOT, "include", "mrt_types.h")]
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.run(["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-w", "-I", CSRC, "-I", os.path.join(ROOT, "include"), "-I", cuda_inc] +
                       srcs + ["-o", exe], check=True)
    return exe


def emul_binned(exe, scene, width, height, spp, depth=32, seed=DEFAULT_SEED, s0=0, s1=0, mode="B", chunk=0, bins=2, crop=None, extra=()):
    from miniraytracer_b200.accfile import read_acc
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        path = f.name
    try:
        r = subprocess.run([exe, "-scene", str(scene), "-width", str(width), "-height", str(height), "-samples", str(spp),
                            "-depth", str(depth), "-seed", str(seed), "-s0", str(s0), "-s1", str(s1), "-mode", mode,
                            "-chunk", str(chunk), "-bins", str(bins), "-assets", ASSETS, "-out", path] + list(extra) +
                           (["-x0", str(crop[0]), "-y0", str(crop[1]), "-x1", str(crop[2]), "-y1", str(crop[3])] if crop else []),
                           check=True, capture_output=True, text=True)
        acc, meta = read_acc(path)
        meta["info"] = json.loads(r.stdout.strip().splitlines()[-1])
        return acc, meta
    finally:
        os.unlink(path)
