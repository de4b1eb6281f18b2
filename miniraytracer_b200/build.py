"""In-tree build of the native code (no JIT cache: the built files travel with the repo snapshot).

  libmrt_b200.so   host scene builder + sm_100a kernels + C ABI (include/mrt_gpu.h)
  mrt_b200         command-line front end (the reference's main() replacement)

Parity build flags: -fmad=false (no FMA contraction), IEEE div/sqrt, no fast-math; host code with
-ffp-contract=off.  Arch: sm_100a only.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "miniraytracer_b200", "csrc")
INC = os.path.join(ROOT, "include")
LIB = os.environ.get("MRT_LIB") or os.path.join(ROOT, "miniraytracer_b200", "libmrt_b200.so")   # MRT_LIB: A/B builds (tools only)
EXE = os.path.join(ROOT, "miniraytracer_b200", "mrt_b200")

HOST_SRCS = ["scene_graph.cpp", "scenes.cpp", "obj_loader.cpp", "scene_dump.cpp", "flatten.cpp", "host_api.cpp"]
CUDA_SRCS = ["render_kernel.cu", "render_variant_cornell.cu", "render_variant_cornell_vol.cu", "render_variant_lists.cu", "render_variant_lists_vol.cu",
             "render_variant_trees.cu", "render_variant_trees_tex.cu", "render_variant_all.cu", "render_variant_full.cu"]
HEADERS = ["scene_graph.h", "trace_core.h", "mrt_libm.h", "gpu_internal.h", "coop_tree.cuh", "render_kernels.cuh", "render_variants.h", "schedule.h", os.path.join(INC, "mrt_types.h"), os.path.join(INC, "mrt_gpu.h")]

EXTRA = os.environ.get("MRT_NVCC_EXTRA", "").split()
NVCC_FLAGS = EXTRA + ["-std=c++20", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-I", CSRC, "-I", INC]
CXX_FLAGS = ["-std=c++20", "-O2", "-ffp-contract=off", "-fPIC", "-Wall", "-I", CSRC, "-I", INC]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("native build failed")
    return r.stdout + r.stderr


def build(force=False, verbose=False):
    """Compile everything that is out of date. Returns the path of the shared library."""
    objdir = os.path.join(ROOT, "build", "obj")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    log = ""
    for src in HOST_SRCS:
        o = os.path.join(objdir, src + ".o")
        s = os.path.join(CSRC, src)
        if force or _stale(o, [s] + hdrs):
            log += _run(["g++"] + CXX_FLAGS + ["-c", s, "-o", o])
        objs.append(o)
    todo = []
    for src in CUDA_SRCS:
        o = os.path.join(objdir, src + ".o")
        s = os.path.join(CSRC, src)
        if force or _stale(o, [s] + hdrs):
            todo.append([_nvcc()] + NVCC_FLAGS + ["-Xptxas", "-v", "-c", s, "-o", o])
        objs.append(o)
    if todo:   # the kernel variants are independent translation units: compile them in parallel
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1)) as ex:
            for out in ex.map(_run, todo):
                log += out
    if force or _stale(LIB, objs):
        log += _run([_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    main_src = os.path.join(CSRC, "main_host.cpp")
    if force or _stale(EXE, [main_src, LIB] + hdrs):
        log += _run(["g++"] + CXX_FLAGS + [main_src, "-o", EXE, "-L", os.path.dirname(LIB), "-lmrt_b200",
                                         "-Wl,-rpath,$ORIGIN", "-lpthread"])
    if verbose and log:
        print(log)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB)
