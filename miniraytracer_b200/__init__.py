"""B200-native renderer for MiniRayTracer's per-pixel path-tracing bounce loop.

Public surface: `api` (ctypes binding of the C ABI in include/mrt_gpu.h), `distributed` (spp sharding +
accumulator all-reduce), `accfile` (accumulator file format / reference finalisation), `build` (in-tree
native build).  The renderer itself is CUDA only; nothing here computes radiance on the CPU.
"""
from . import accfile, build  # noqa: F401

__all__ = ["accfile", "api", "build", "distributed"]
