"""ctypes binding of the C ABI (include/mrt_gpu.h) -- the only way Python reaches the renderer.

There is no CPU rendering path: `Renderer` raises if the shared library or a CUDA device is missing.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None


class F4(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


class Camera(C.Structure):
    _fields_ = [(n, C.c_float * 3) for n in ("origin", "u", "v", "w", "llcorner", "horz", "vert")] + \
               [("lens_radius", C.c_float), ("time0", C.c_float), ("time1", C.c_float)]


class SceneDesc(C.Structure):
    _fields_ = [
        ("root", C.c_uint32), ("n_lights", C.c_uint32), ("lights", C.POINTER(C.c_uint32)),
        ("sky", C.c_uint32), ("stack_words", C.c_uint32), ("features", C.c_uint32), ("stack_words_coop", C.c_uint32),
        ("camera", Camera),
        ("sphere", C.POINTER(F4)), ("n_sphere", C.c_uint32),
        ("rect", C.POINTER(F4)), ("n_rect", C.c_uint32),
        ("list", C.POINTER(F4)), ("n_list", C.c_uint32),
        ("child", C.POINTER(C.c_uint32)), ("n_child", C.c_uint32),
        ("bvh", C.POINTER(F4)), ("n_bvh", C.c_uint32),
        ("node2", C.POINTER(F4)), ("n_node2", C.c_uint32),
        ("trileaf", C.POINTER(C.c_uint32)), ("n_trileaf", C.c_uint32),
        ("tri", C.POINTER(F4)), ("n_tri", C.c_uint32),
        ("trin", C.POINTER(F4)),
        ("xlate", C.POINTER(F4)), ("n_xlate", C.c_uint32),
        ("rot", C.POINTER(F4)), ("n_rot", C.c_uint32),
        ("vol", C.POINTER(F4)), ("n_vol", C.c_uint32),
        ("mat", C.POINTER(F4)), ("n_mat", C.c_uint32),
        ("tex", C.POINTER(F4)), ("n_tex", C.c_uint32),
        ("perlin_vec", C.POINTER(F4)), ("perlin_perm", C.POINTER(C.c_int32)),
        ("image", C.POINTER(C.c_uint8)), ("n_image_bytes", C.c_uint64),
    ]


class RenderParams(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("samples", C.c_uint32),
                ("sample_begin", C.c_uint32), ("sample_end", C.c_uint32), ("max_bounces", C.c_uint32),
                ("seed", C.c_uint64), ("max_luminance", C.c_float), ("flags", C.c_uint32),
                ("crop_x0", C.c_uint32), ("crop_y0", C.c_uint32), ("crop_x1", C.c_uint32), ("crop_y1", C.c_uint32)]


class Tuning(C.Structure):
    """MrtTuning: scheduling knobs (all zero = the measured defaults)."""
    _fields_ = [("mode", C.c_uint32), ("bins", C.c_uint32), ("min_blocks", C.c_uint32), ("chunk_pixels", C.c_uint32),
                ("variant_all", C.c_uint32), ("z_order", C.c_uint32), ("coop_trees", C.c_uint32), ("coop_leaf_batch", C.c_uint32),
                ("chunk_paths", C.c_uint32), ("tail_tasks", C.c_uint32), ("blocks_per_sm", C.c_uint32), ("reserved", C.c_uint32 * 5)]


MODE_AUTO, MODE_PER_LANE, MODE_PER_WARP, MODE_BINNED = 0, 1, 2, 3


class Params(C.Structure):
    _fields_ = [("window_width", C.c_uint32), ("window_height", C.c_uint32),
                ("buffer_width", C.c_uint32), ("buffer_height", C.c_uint32),
                ("samples_per_pixel", C.c_uint32), ("tile_size", C.c_uint32), ("num_threads", C.c_uint32),
                ("max_bounces", C.c_uint32), ("scene_select", C.c_uint32), ("threading_mode", C.c_uint32),
                ("max_luminance", C.c_float), ("delay", C.c_uint32), ("num_gpus", C.c_uint32),
                ("seed", C.c_uint64), ("out_path", C.c_char * 512), ("asset_dir", C.c_char * 512)]


class DeviceInfo(C.Structure):
    _fields_ = [("device", C.c_int), ("sm_count", C.c_int), ("clock_khz", C.c_int), ("cc_major", C.c_int),
                ("cc_minor", C.c_int), ("total_mem", C.c_uint64), ("name", C.c_char * 128)]


class RenderStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("paths", C.c_uint64), ("nonfinite", C.c_uint64), ("warp_iterations", C.c_uint64), ("kernel_ms", C.c_float),
                ("grid", C.c_uint32), ("block", C.c_uint32), ("smem_bytes", C.c_uint32), ("mode", C.c_uint32), ("coop_trees", C.c_uint32),
                ("coop_node_steps", C.c_uint64), ("coop_node_items", C.c_uint64), ("coop_leaf_steps", C.c_uint64), ("coop_leaf_items", C.c_uint64),
                ("warp_time_sum_ns", C.c_uint64), ("warp_span_ns", C.c_uint64), ("first_exit_ns", C.c_uint64), ("warps", C.c_uint32), ("reserved", C.c_uint32), ("stage_sum_ns", C.c_uint64)]


MRT_RENDER_ACCUMULATE = 1
MRT_RENDER_CONTINUE = 2
SCENE_EXTRA_TRIANGLES = 0x200   # MRT_SCENE_EXTRA_TRIANGLES: two triangle_scene_objects in the Cornell box
SCENE_ALL_LIGHTS = 0x100   # MRT_SCENE_ALL_LIGHTS: light list with both allocated entries (ceiling light + glass sphere)
DEFAULT_SEED = 11350390909718046443  # main.cpp:302

# every symbol include/mrt_gpu.h declares
EXPORTS = [
    "mrt_last_error", "mrt_params_default", "mrt_params_parse", "mrt_scene_create", "mrt_scene_desc",
    "mrt_scene_dump", "mrt_scene_save", "mrt_scene_load", "mrt_scene_free", "mrt_gpu_init", "mrt_gpu_scene_upload", "mrt_gpu_set_tuning", "mrt_gpu_set_stream",
    "mrt_gpu_bind_accumulator", "mrt_gpu_render_async", "mrt_gpu_poll", "mrt_gpu_wait", "mrt_gpu_stats",
    "mrt_gpu_finalize_device", "mrt_gpu_running_mean_update", "mrt_gpu_render_running_mean", "mrt_gpu_readback", "mrt_gpu_tonemap", "mrt_gpu_tonemap_device", "mrt_gpu_reduce_finalize", "mrt_gpu_cancel", "mrt_gpu_destroy",
]


class MrtError(RuntimeError):
    pass


def lib_path():
    return _build.LIB


def load(build_if_missing=True):
    """Load libmrt_b200.so (building it in-tree first if the sources are newer)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if build_if_missing and os.path.isdir(_build.CSRC) and os.environ.get("MRT_NO_BUILD") != "1":
        try:
            _build.build()
        except Exception:
            if not os.path.exists(_build.LIB):
                raise
    if not os.path.exists(_build.LIB):
        raise MrtError(f"{_build.LIB} is missing: run `python -m miniraytracer_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(_build.LIB)
    vp = C.c_void_p
    lib.mrt_last_error.restype = C.c_char_p
    lib.mrt_params_default.argtypes = [C.POINTER(Params)]
    lib.mrt_params_default.restype = None
    lib.mrt_params_parse.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(Params)]
    lib.mrt_scene_create.argtypes = [C.c_uint32, C.c_float, C.c_char_p, C.POINTER(vp)]
    lib.mrt_scene_desc.argtypes = [vp]
    lib.mrt_scene_desc.restype = C.POINTER(SceneDesc)
    lib.mrt_scene_dump.argtypes = [vp, C.c_char_p]
    lib.mrt_scene_save.argtypes = [vp, C.c_char_p]
    lib.mrt_scene_load.argtypes = [C.c_char_p, C.POINTER(vp)]
    lib.mrt_scene_free.argtypes = [vp]
    lib.mrt_scene_free.restype = None
    lib.mrt_gpu_init.argtypes = [C.c_int, C.POINTER(DeviceInfo)]
    lib.mrt_gpu_scene_upload.argtypes = [C.POINTER(SceneDesc), C.POINTER(vp)]
    lib.mrt_gpu_set_tuning.argtypes = [vp, C.POINTER(Tuning)]
    lib.mrt_gpu_set_stream.argtypes = [vp, vp]
    lib.mrt_gpu_bind_accumulator.argtypes = [vp, vp, C.c_uint32, C.c_uint32]
    lib.mrt_gpu_render_async.argtypes = [vp, C.POINTER(RenderParams)]
    lib.mrt_gpu_poll.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_uint64)]
    lib.mrt_gpu_wait.argtypes = [vp]
    lib.mrt_gpu_stats.argtypes = [vp, C.POINTER(RenderStats)]
    lib.mrt_gpu_finalize_device.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint32, C.c_float]
    lib.mrt_gpu_readback.argtypes = [vp, vp, C.c_int]
    lib.mrt_gpu_running_mean_update.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float]
    lib.mrt_gpu_render_running_mean.argtypes = [vp, C.POINTER(RenderParams), vp]
    lib.mrt_gpu_tonemap.argtypes = [vp, vp]
    lib.mrt_gpu_tonemap_device.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint32]
    lib.mrt_gpu_reduce_finalize.argtypes = [C.POINTER(vp), C.c_int, C.c_float, vp, vp]
    lib.mrt_gpu_cancel.argtypes = [vp]
    lib.mrt_gpu_destroy.argtypes = [vp]
    lib.mrt_gpu_destroy.restype = None
    _LIB = lib
    return lib


def _check(rc):
    if rc != 0:
        raise MrtError(f"mrt error {rc}: {load().mrt_last_error().decode(errors='replace')}")


def default_asset_dir():
    return os.environ.get("MRT_ASSET_DIR", os.path.join(_build.ROOT, "assets"))


def parse_args(argv):
    """mrt_params_parse: returns (Params, help_requested)."""
    lib = load()
    arr = (C.c_char_p * (len(argv) + 1))(*([b"mrt"] + [a.encode() for a in argv]))
    p = Params()
    rc = lib.mrt_params_parse(len(argv) + 1, arr, C.byref(p))
    return p, bool(rc)


def grid_samples(spp):
    """N = floor(sqrt(spp))^2 (main.cpp:319-320)."""
    sq = int(np.sqrt(np.float32(spp)))
    return sq * sq


class HostScene:
    """Host-built + flattened scene (mrt_scene_create). Usable without a GPU."""

    def __init__(self, scene, width, height, asset_dir=None, path=None):
        lib = load()
        self._lib = lib
        self._h = C.c_void_p()
        if path is not None:   # flattened scene file written by save()
            _check(lib.mrt_scene_load(str(path).encode(), C.byref(self._h)))
            self.scene = None
            return
        aspect = np.float32(width) / np.float32(height)
        _check(lib.mrt_scene_create(int(scene), C.c_float(aspect), (asset_dir or default_asset_dir()).encode(),
                                    C.byref(self._h)))
        self.scene = int(scene)   # may carry SCENE_ALL_LIGHTS

    @classmethod
    def load(cls, path):
        return cls(None, 0, 0, path=path)

    def save(self, path):
        _check(self._lib.mrt_scene_save(self._h, str(path).encode()))

    @property
    def desc(self):
        return self._lib.mrt_scene_desc(self._h)

    def dump(self, path):
        _check(self._lib.mrt_scene_dump(self._h, str(path).encode()))

    def close(self):
        if self._h:
            self._lib.mrt_scene_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Renderer:
    """A scene resident on one GPU (mrt_gpu_scene_upload) + render / readback calls."""

    def __init__(self, host_scene, device=0, tuning=None):
        lib = load()
        self._lib = lib
        self.info = DeviceInfo()
        _check(lib.mrt_gpu_init(int(device), C.byref(self.info)))
        self.device = int(device)
        self._h = C.c_void_p()
        _check(lib.mrt_gpu_scene_upload(host_scene.desc, C.byref(self._h)))
        self._size = None
        if tuning:
            self.set_tuning(**tuning)

    def set_tuning(self, **kw):
        """mrt_gpu_set_tuning: keyword = MrtTuning field (mode, bins, min_blocks, chunk_pixels, variant_all, z_order,
        coop_trees); no keywords = back to the defaults."""
        t = Tuning()
        for k, v in kw.items():
            if k not in dict(Tuning._fields_) or k == "reserved":
                raise MrtError(f"unknown tuning field {k!r}")
            setattr(t, k, int(v))
        _check(self._lib.mrt_gpu_set_tuning(self._h, C.byref(t)))

    def set_stream(self, cuda_stream_ptr):
        _check(self._lib.mrt_gpu_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def bind_accumulator(self, device_ptr, width, height):
        _check(self._lib.mrt_gpu_bind_accumulator(self._h, C.c_void_p(device_ptr), width, height))

    def render_async(self, width, height, spp, depth=32, seed=DEFAULT_SEED, sample_begin=0, sample_end=None,
                     max_luminance=1000.0, accumulate=False, crop=None):
        """crop = (x0, y0, x1, y1): render only that window of the width x height frame; the accumulator (and every
        readback) then has the window's size."""
        n = grid_samples(spp)
        x0, y0, x1, y1 = crop if crop else (0, 0, 0, 0)
        p = RenderParams(width, height, n, sample_begin, n if sample_end is None else sample_end, depth, seed,
                         max_luminance, MRT_RENDER_ACCUMULATE if accumulate else 0, x0, y0, x1, y1)
        _check(self._lib.mrt_gpu_render_async(self._h, C.byref(p)))
        self._size = (x1 - x0, y1 - y0) if crop else (width, height)
        return p

    def render_running_mean(self, width, height, spp, depth=32, seed=DEFAULT_SEED, sample_begin=0, sample_end=None, max_luminance=1000.0):
        """mrt_gpu_render_running_mean: draw2's sample-major passes with the per-pass clamped running mean; returns mean[h, w, 4]."""
        n = grid_samples(spp)
        p = RenderParams(width, height, n, sample_begin, n if sample_end is None else sample_end, depth, seed, max_luminance, 0, 0, 0, 0, 0)
        out = np.empty((height, width, 4), dtype=np.float32)
        _check(self._lib.mrt_gpu_render_running_mean(self._h, C.byref(p), out.ctypes.data_as(C.c_void_p)))
        self._size = (width, height)
        return out

    def poll(self):
        pct, rays = C.c_float(), C.c_uint64()
        _check(self._lib.mrt_gpu_poll(self._h, C.byref(pct), C.byref(rays)))
        return pct.value, rays.value

    def wait(self):
        _check(self._lib.mrt_gpu_wait(self._h))

    def stats(self):
        st = RenderStats()
        _check(self._lib.mrt_gpu_stats(self._h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in RenderStats._fields_}

    def finalize_device(self, acc_ptr, out_ptr, width, height, max_luminance=1000.0):
        _check(self._lib.mrt_gpu_finalize_device(self._h, C.c_void_p(acc_ptr), C.c_void_p(out_ptr), width, height,
                                                 C.c_float(max_luminance)))

    def readback(self, finalize=False, out=None):
        if self._size is None:
            raise MrtError("readback before any render")
        w, h = self._size
        if out is None:
            out = np.empty((h, w, 4), dtype=np.float32)
        _check(self._lib.mrt_gpu_readback(self._h, out.ctypes.data_as(C.c_void_p), 1 if finalize else 0))
        return out

    def readback_into(self, host_ptr, finalize=False):
        _check(self._lib.mrt_gpu_readback(self._h, C.c_void_p(host_ptr), 1 if finalize else 0))

    def tonemap(self):
        w, h = self._size
        out = np.empty((h, w), dtype=np.uint32)
        _check(self._lib.mrt_gpu_tonemap(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def tonemap_device(self, img_ptr, argb_ptr, width, height):
        _check(self._lib.mrt_gpu_tonemap_device(self._h, C.c_void_p(img_ptr), C.c_void_p(argb_ptr), width, height))

    def cancel(self):
        _check(self._lib.mrt_gpu_cancel(self._h))

    def close(self):
        if self._h:
            self._lib.mrt_gpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def reduce_finalize(renderers, max_luminance=1000.0, tonemap=False):
    """mrt_gpu_reduce_finalize: sum the accumulators of renderers that rendered the same frame (one sample slice each, one GPU
    each), finalise over NVLink peer memory; returns the finalised float image (and the ARGB tone map if asked)."""
    lib = load()
    w, h = renderers[0]._size
    arr = (C.c_void_p * len(renderers))(*[r._h.value for r in renderers])
    img = np.empty((h, w, 4), dtype=np.float32)
    argb = np.empty((h, w), dtype=np.uint32) if tonemap else None
    _check(lib.mrt_gpu_reduce_finalize(arr, len(renderers), C.c_float(max_luminance), img.ctypes.data_as(C.c_void_p),
                                       argb.ctypes.data_as(C.c_void_p) if tonemap else None))
    return (img, argb) if tonemap else img


def render(scene, width, height, spp, depth=32, seed=DEFAULT_SEED, device=0, asset_dir=None, finalize=False, tuning=None, **kw):
    """One-call convenience: build + upload the scene, render all samples, return (image, stats)."""
    hs = HostScene(scene, width, height, asset_dir)
    r = Renderer(hs, device, tuning)
    try:
        r.render_async(width, height, spp, depth, seed, **kw)
        st = r.stats()
        return r.readback(finalize=finalize), st
    finally:
        r.close()
        hs.close()
