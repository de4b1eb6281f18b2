"""The C-ABI library loads without a GPU and exports every symbol include/mrt_gpu.h declares;
parameter parsing mirrors cmdline_parser.cpp."""
import ctypes
import os
import re

import pytest

from miniraytracer_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mrt_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mrt_[a-z0-9_]+)\s*\(", src)))


def test_all_declared_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mrt_gpu.h but not exported"
    assert sorted(api.EXPORTS) == names


def test_struct_sizes_match_header(lib):
    # cross-check the ctypes mirrors against the C compiler's view of the header
    import subprocess, tempfile, textwrap
    code = textwrap.dedent("""
        #include <stdio.h>
        #include "mrt_gpu.h"
        int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(MrtSceneDesc), sizeof(MrtRenderParams), sizeof(MrtParams),
                         sizeof(MrtDeviceInfo), sizeof(MrtRenderStats), sizeof(MrtCamera), sizeof(MrtF4)); return 0; }
    """)
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(code)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)   # header is plain C
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    mine = [ctypes.sizeof(t) for t in (api.SceneDesc, api.RenderParams, api.Params, api.DeviceInfo, api.RenderStats, api.Camera, api.F4)]
    assert sizes == mine


def test_params_defaults_and_parse():
    p, helped = api.parse_args([])
    assert not helped
    # cmdline_parser.h:5-18
    assert (p.window_width, p.window_height, p.samples_per_pixel, p.tile_size, p.max_bounces) == (500, 500, 128, 32, 32)
    assert p.scene_select == 8 and p.threading_mode == 1 and p.max_luminance == 1000.0 and p.delay == 0
    p, _ = api.parse_args("-scene 5 -width 1920 -height 1080 -samples 1024 -depth 32 -delay -gpus 8 -seed 7".split())
    assert (p.scene_select, p.buffer_width, p.buffer_height, p.samples_per_pixel, p.max_bounces) == (5, 1920, 1080, 1024, 32)
    assert p.delay == 1 and p.num_gpus == 8 and p.seed == 7
    # out-of-range values are rejected with a warning and keep the default (cmdline_parser.cpp:52-55)
    p, _ = api.parse_args("-scene 9 -mode 2 -width 0".split())
    assert p.scene_select == 8 and p.threading_mode == 1 and p.buffer_width == 500
    _, helped = api.parse_args(["--help"])
    assert helped


def test_errors_are_status_codes(lib):
    h = ctypes.c_void_p()
    assert lib.mrt_scene_create(99, ctypes.c_float(1.0), b"assets", ctypes.byref(h)) != 0
    assert b"scene" in lib.mrt_last_error()
    assert lib.mrt_gpu_render_async(None, None) != 0
    assert lib.mrt_gpu_readback(None, None, 0) != 0


def test_grid_samples():
    assert api.grid_samples(16) == 16 and api.grid_samples(128) == 121 and api.grid_samples(1024) == 1024
    assert api.grid_samples(4096) == 4096 and api.grid_samples(1) == 1
