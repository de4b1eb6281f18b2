"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle (the reference's own
renderer with identical per-(pixel, sample) RNG streams).

Bar (BASELINE.json north_star): a low-spp run with identical RNG streams must match per pixel within
1e-4 relative on at least 99.9% of pixels.  The arithmetic is IEEE float32 in the reference's operation
order; only libm calls (sinf/cosf/atan2f/asinf/logf/powf) differ by the library's ulp error."""
import os

import numpy as np
import pytest

import oracle_util
from miniraytracer_b200 import accfile, api

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
REL_TOL = 1e-4          # per-pixel relative tolerance stated by north_star
MIN_FRAC = 0.999        # fraction of pixels that must agree


def _gpu_render(scene, w, h, spp, depth=32, seed=api.DEFAULT_SEED, tuning=None, **kw):
    hs = api.HostScene(scene, w, h)
    r = api.Renderer(hs, 0, tuning)
    try:
        r.render_async(w, h, spp, depth, seed, **kw)
        st = r.stats()
        return r.readback(), st
    finally:
        r.close()
        hs.close()


def _check(acc, ref, ref_rays=None, st=None):
    res = accfile.compare(accfile.finalize(acc), accfile.finalize(ref), rel=REL_TOL)
    dropped_diff = int((acc[..., 3] != ref[..., 3]).sum())
    assert res["frac_ok"] >= MIN_FRAC, (res, dropped_diff)
    if ref_rays is not None and st is not None:
        # trace() call count (G_rayCounter): equal up to the rare paths whose discrete decisions flip on an ulp
        assert abs(int(st["rays"]) - int(ref_rays)) <= 2e-3 * ref_rays, (st["rays"], ref_rays)
    return res


@pytest.mark.parametrize("scene", range(9))
def test_golden_fixtures(scene):
    g = np.load(os.path.join(GOLDEN, f"golden_scene{scene}.npz"))
    acc, st = _gpu_render(scene, int(g["width"]), int(g["height"]), int(g["spp"]), int(g["depth"]))
    _check(acc, g["acc"], int(g["rays"]), st)


needs_ref = pytest.mark.skipif(not oracle_util.have_ref(), reason="oracle/_ref/mrt_ref not on this box")


@pytest.mark.parametrize("scene", [5, 7])
def test_sphere_light_golden(scene):
    """Light list {ceiling light, glass sphere} (MRT_SCENE_ALL_LIGHTS; oracle: `-lights all` = the reference's list with its count
    set to the 2 entries it allocates, scene.cpp:326-329,456-459): sphere::pdf_value / pdf_generate (sphere.cpp:63-79) and
    random_towards_sphere (pcg.cpp:125-133), dead code in the stock scenes, against the reference."""
    g = np.load(os.path.join(GOLDEN, f"golden_all_lights_scene{scene}.npz"))
    acc, st = _gpu_render(scene | api.SCENE_ALL_LIGHTS, int(g["width"]), int(g["height"]), int(g["spp"]), int(g["depth"]))
    _check(acc, g["acc"], int(g["rays"]), st)
    stock = np.load(os.path.join(GOLDEN, f"golden_scene{scene}.npz"))
    assert int(g["rays"]) != int(stock["rays"]) * 4      # (the stock golden has 4 spp) the second light changes the paths


@needs_ref
@pytest.mark.parametrize("scene,w,h,spp", [(5, 240, 135, 36), (7, 160, 90, 16)])
def test_sphere_light_parity_vs_oracle(scene, w, h, spp):
    ref, meta = oracle_util.ref_render(scene, w, h, spp, all_lights=True)
    acc, st = _gpu_render(scene | api.SCENE_ALL_LIGHTS, w, h, spp)
    _check(acc, ref, meta["rays"], st)

# BASELINE.json configs: C1 exactly; C2..C5 at the config's scene/aspect/depth with the resolution and spp the
# CPU oracle finishes in seconds (per-pixel parity needs identical streams, not convergence)
CONFIGS = [
    ("C1", 0, 500, 500, 16),
    ("C2", 5, 480, 270, 16),
    ("C3", 6, 480, 270, 16),
    ("C4", 7, 480, 270, 16),
    ("C5", 8, 480, 270, 16),
]


@needs_ref
@pytest.mark.parametrize("name,scene,w,h,spp", CONFIGS)
def test_config_parity_vs_oracle(name, scene, w, h, spp):
    ref, meta = oracle_util.ref_render(scene, w, h, spp)
    acc, st = _gpu_render(scene, w, h, spp)
    res = _check(acc, ref, meta["rays"], st)
    print(name, res, "gpu rays", st["rays"], "ref rays", meta["rays"], "kernel ms", st["kernel_ms"])


@needs_ref
@pytest.mark.parametrize("name,scene,w,h,spp", CONFIGS)
def test_config_parity_vs_reference_with_host_libm(name, scene, w, h, spp):
    """The bar also holds against the reference AS BUILT: the same oracle binary with MRT_ORACLE_LIBM=host, i.e. this box's glibc
    sinf / cosf / atan2f / asinf / logf / powf instead of the correctly rounded ones both sides use by default.  A 1-ulp libm
    difference in a sampled direction flips a few paths, so pixels are no longer bit-identical, but >= 99.9 % stay within 1e-4
    (profiles/r2_libm_pass_rates.jsonl: 99.96 - 100 %)."""
    ref, meta = oracle_util.ref_render(scene, w, h, spp, host_libm=True)
    acc, st = _gpu_render(scene, w, h, spp)
    res = accfile.compare(accfile.finalize(acc), accfile.finalize(ref), rel=REL_TOL)
    assert res["frac_ok"] >= MIN_FRAC, res
    assert abs(int(st["rays"]) - int(meta["rays"])) <= 1e-4 * meta["rays"]


@needs_ref
@pytest.mark.parametrize("scene,w,h,spp", [(5, 240, 136, 64), (0, 200, 200, 36)])
def test_pixel_per_warp_mode(scene, w, h, spp):
    # MrtTuning.mode = MRT_MODE_PER_WARP selects the plain pixel-per-warp kernel (mode W: lanes keep their paths, lane
    # sums combined by a shuffle tree); the default is mode B (test below)
    ref, meta = oracle_util.ref_render(scene, w, h, spp)
    acc, st = _gpu_render(scene, w, h, spp, tuning=dict(mode=api.MODE_PER_WARP))
    assert st["mode"] == api.MODE_PER_WARP
    _check(acc, ref, meta["rays"], st)


@pytest.mark.parametrize("mode", [api.MODE_PER_LANE, api.MODE_BINNED])
def test_modes_and_slices_agree(mode):
    # 64 spp in one launch == four accumulated 16-sample slices, in the pixel-per-lane kernel (mode P) and in mode B
    slice_mode = mode
    tuning = dict(mode=mode)
    w, h, spp = 160, 90, 64
    full, st = _gpu_render(5, w, h, spp, tuning=tuning)
    hs = api.HostScene(5, w, h)
    r = api.Renderer(hs, 0, tuning)
    for i in range(4):
        r.render_async(w, h, spp, sample_begin=16 * i, sample_end=16 * (i + 1), accumulate=(i > 0))
        assert r.stats()["mode"] == slice_mode
    parts = r.readback()
    r.close(); hs.close()
    np.testing.assert_array_equal(parts[..., 3], full[..., 3])
    res = accfile.compare(accfile.finalize(parts), accfile.finalize(full), rel=1e-5)
    assert res["n_bad"] == 0, res


def test_deterministic_run_to_run():
    a, _ = _gpu_render(7, 160, 90, 36)
    b, _ = _gpu_render(7, 160, 90, 36)
    np.testing.assert_array_equal(a, b)
    c, _ = _gpu_render(0, 128, 128, 16)
    d, _ = _gpu_render(0, 128, 128, 16)
    np.testing.assert_array_equal(c, d)


def test_seed_changes_image_and_depth_zero():
    a, _ = _gpu_render(0, 96, 96, 4, seed=1)
    b, _ = _gpu_render(0, 96, 96, 4, seed=2)
    assert np.abs(a - b).max() > 0
    z, st = _gpu_render(5, 96, 54, 4, depth=0)    # depth 0: only directly visible emitters
    assert st["rays"] == 96 * 54 * 4


def test_ragged_sizes_and_single_sample():
    # odd sizes, fewer pixels than lanes, one sample
    for (w, h, spp) in [(1, 1, 1), (7, 3, 1), (33, 5, 4), (5, 1, 49)]:
        acc, st = _gpu_render(5, w, h, spp)
        assert acc.shape == (h, w, 4)
        assert np.all(acc[..., 3] <= api.grid_samples(spp))
        assert st["paths"] == w * h * api.grid_samples(spp)


@needs_ref
def test_finalize_matches_reference_accumulate():
    # mean + luminance clamp (main.cpp:168-173) on the device vs the numpy restatement
    w, h = 120, 68
    hs = api.HostScene(5, w, h)
    r = api.Renderer(hs, 0)
    r.render_async(w, h, 16, max_luminance=0.5)
    raw = r.readback()
    fin = r.readback(finalize=True)
    argb = r.tonemap()
    r.close(); hs.close()
    want = accfile.finalize(raw, 0.5)
    np.testing.assert_allclose(fin[..., :3], want, rtol=1e-6, atol=1e-7)
    assert argb.shape == (h, w) and argb.max() > 0      # values: test_tonemap_matches_reference


def test_bad_arguments():
    hs = api.HostScene(5, 64, 36)
    r = api.Renderer(hs, 0)
    with pytest.raises(api.MrtError):
        r.render_async(64, 36, 16, sample_begin=8, sample_end=4)
    with pytest.raises(api.MrtError):
        r.render_async(0, 36, 16)
    with pytest.raises(api.MrtError):
        r.readback()          # nothing rendered yet
    r.close(); hs.close()


def test_poll_and_cancel():
    w, h = 640, 360
    hs = api.HostScene(5, w, h)
    r = api.Renderer(hs, 0)
    r.render_async(w, h, 4096)
    pct, _ = r.poll()
    assert 0.0 <= pct <= 100.0
    r.cancel()
    r.wait()
    st = r.stats()
    assert st["rays"] < w * h * 4096 * 1.5   # stopped early (a full render is ~2.1 rays per path)
    r.render_async(w, h, 4)                  # the handle is reusable after a cancel
    r.wait()
    pct, rays = r.poll()
    assert pct == 100.0 and rays > 0
    r.close(); hs.close()


@needs_ref
@pytest.mark.parametrize("binned", [None, 1, 2, 3])
@pytest.mark.parametrize("scene,w,h,spp", [(5, 160, 90, 64), (6, 128, 72, 36), (7, 96, 54, 36), (8, 96, 54, 36), (0, 96, 96, 49)])
def test_binned_pool_renderer_parity(scene, w, h, spp, binned):
    """Mode B (paths parked in a per-warp pool and regrouped by a ray classifier between segments; MrtTuning.bins: one bin,
    classifier bins, + pending-weight bit) runs the same per-path arithmetic as the other modes: same ray count,
    accumulators equal up to the order of the per-pixel sum."""
    ref, meta = oracle_util.ref_render(scene, w, h, spp)
    tuning = None if binned is None else dict(bins=binned)   # None: the default
    acc, st = _gpu_render(scene, w, h, spp, tuning=tuning)
    assert st["mode"] == api.MODE_BINNED
    _check(acc, ref, meta["rays"], st)
    acc2, _ = _gpu_render(scene, w, h, spp, tuning=tuning)
    assert np.array_equal(acc, acc2), "binned schedule must be reproducible run to run"


def test_progressive_passes_single_gpu():
    """distributed.ProgressiveReducer on one GPU: sample-major passes accumulated into a bound torch accumulator with
    the per-pass snapshot taken on a side stream; the last preview equals a one-shot render of all samples.  The renderer
    runs on a NON-default torch stream that the reducer is told about (renderer=, stream=): snapshot and render are ordered
    by events on that stream, not on whatever stream happens to be current."""
    import torch
    from miniraytracer_b200 import distributed as mdist
    w, h, spp, passes = 160, 90, 64, 4
    full, _ = _gpu_render(5, w, h, spp)
    hs = api.HostScene(5, w, h)
    r = api.Renderer(hs, 0)
    render_stream = torch.cuda.Stream(device="cuda:0")
    acc = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda:0")
    torch.cuda.synchronize()
    r.bind_accumulator(acc.data_ptr(), w, h)
    counts = []

    def render_pass(b, e, out):
        r.render_async(w, h, spp, sample_begin=b, sample_end=e, accumulate=True)

    red = mdist.ProgressiveReducer(acc, render_pass, preview=lambda p, buf: counts.append(float(buf[..., 3].max().item())),
                                   stream=render_stream, renderer=r)
    final = red.run(mdist.progressive_schedule(spp, 0, 1, passes))
    torch.cuda.synchronize()
    got = final.cpu().numpy()
    empty = mdist.ProgressiveReducer(acc, render_pass, stream=render_stream, renderer=r).run([])
    assert empty is not None and torch.equal(empty, acc)
    r.close(); hs.close()
    assert counts == [16.0, 32.0, 48.0, 64.0]
    np.testing.assert_array_equal(got[..., 3], full[..., 3])
    res = accfile.compare(accfile.finalize(got), accfile.finalize(full), rel=1e-5)
    assert res["n_bad"] == 0, res


def test_few_samples_are_added_in_sample_order_and_warp_times_are_reported():
    """Launches with at most 128 samples per pixel run the SEQ instantiation of mode B: a pixel's samples are added one after
    the other in sample order (the reference's loop, main.cpp:154-166), so the accumulator equals the float32 running sum of
    one-sample launches; more samples per pixel use lane-strided partial sums (same samples, sum within rounding).  Also the
    mode-B timing counters of MrtRenderStats."""
    w, h, spp = 40, 22, 121
    acc, st = _gpu_render(5, w, h, spp)
    assert st["mode"] == api.MODE_BINNED
    run = np.zeros_like(acc)
    hs = api.HostScene(5, w, h)
    r = api.Renderer(hs, 0)
    try:
        for s in range(spp):
            r.render_async(w, h, spp, 32, api.DEFAULT_SEED, sample_begin=s, sample_end=s + 1)
            run = (run + r.readback()).astype(np.float32)
        np.testing.assert_array_equal(acc, run)
        r.render_async(w, h, 169, 32, api.DEFAULT_SEED)     # > 128 samples per pixel
        big, st2 = r.readback(), r.stats()
    finally:
        r.close()
        hs.close()
    np.testing.assert_array_equal(big[..., 3], np.float32(169))
    assert st2["warps"] == st2["grid"] * st2["block"] // 32 and st2["warp_span_ns"] > 0
    assert 0 < st2["warp_time_sum_ns"] <= st2["warps"] * st2["warp_span_ns"]
    assert 0 < st2["stage_sum_ns"] < st2["warp_time_sum_ns"] and st2["first_exit_ns"] <= st2["warp_span_ns"]
    assert abs(st2["warp_span_ns"] * 1e-6 - st2["kernel_ms"]) < 0.5 * st2["kernel_ms"] + 0.05


def test_binned_result_does_not_depend_on_the_schedule():
    """Mode B sums the finished samples of a pixel in item order from its staging array, so the accumulator is
    bit-identical whatever the bins, the chunk size or the launch bounds (i.e. whichever lane ran which path)."""
    ref = None
    for tuning in (dict(bins=1), dict(bins=2), dict(bins=3), dict(bins=2, chunk_pixels=3), dict(bins=2, min_blocks=8),
                   dict(chunk_paths=512), dict(chunk_paths=1024, tail_tasks=8)):
        acc, st = _gpu_render(6, 128, 72, 64, tuning=tuning)
        assert st["mode"] == api.MODE_BINNED
        if ref is None: ref = acc
        else: np.testing.assert_array_equal(acc, ref)


def test_deep_paths_fall_back_to_mode_w():
    """A bounce limit above 255 does not fit the parked path's 8-bit depth field: such launches use mode W."""
    acc, st = _gpu_render(5, 96, 54, 36, depth=300)
    assert st["mode"] == api.MODE_PER_WARP
    ref, st2 = _gpu_render(5, 96, 54, 36, depth=255)
    assert st2["mode"] == api.MODE_BINNED
    # in the Cornell box paths practically never reach 255 bounces, so the two images agree sample for sample
    np.testing.assert_array_equal(acc[..., 3], ref[..., 3])
    res = accfile.compare(accfile.finalize(acc), accfile.finalize(ref), rel=1e-4)
    assert res["frac_ok"] >= MIN_FRAC, res


@needs_ref
@pytest.mark.parametrize("mode", [api.MODE_AUTO, api.MODE_PER_WARP, api.MODE_PER_LANE])
def test_ragged_frames_and_slices_vs_oracle(mode):
    """Edge shapes against the oracle in mode B (default) and modes W / P: fewer pixels than lanes, sample counts that
    are not multiples of 32, a sample slice that starts and ends inside the grid, a scene with a BVH."""
    for scene, w, h, spp, s0, s1 in [(5, 7, 3, 1, 0, 1), (5, 33, 5, 4, 0, 4), (6, 5, 1, 49, 0, 49), (0, 31, 9, 36, 0, 36),
                                      (5, 40, 22, 16, 5, 12), (8, 24, 13, 100, 37, 90)]:
        ref, meta = oracle_util.ref_render(scene, w, h, spp, s0=s0, s1=s1)
        acc, st = _gpu_render(scene, w, h, spp, sample_begin=s0, sample_end=s1, tuning=dict(mode=mode))
        assert st["paths"] == w * h * (s1 - s0)
        assert st["rays"] == meta["rays"], (scene, w, h, spp, st["rays"], meta["rays"])
        np.testing.assert_array_equal(acc[..., 3], ref[..., 3])
        res = accfile.compare(accfile.finalize(acc), accfile.finalize(ref), rel=REL_TOL)
        assert res["n_bad"] == 0, (scene, w, h, spp, res)


# ---- parity at BASELINE.json's TRUE configuration sizes, through crop windows (the oracle traces only the window; PCG32
# stream ids ((y*W+x)*N+s, pcg.cpp:28-35, main.cpp:156-157) and sub-pixel positions are those of the full frame).  C4 and C5
# have stream ids above 2^32 (2^34.9 in the C5_upper window; the frame's corners are outside the Cornell box, i.e. black).  Golden windows are committed
# (tests/golden/make_golden_windows.py), so this also runs where the oracle binary is absent.
WINDOWS = [
    # name, scene, W, H, spp, (x0, y0, x1, y1)
    ("C2_center", 5, 1920, 1080, 1024, (928, 508, 992, 572)),
    ("C2_upper", 5, 1920, 1080, 1024, (1300, 1000, 1364, 1064)),
    ("C3_smoke", 6, 1920, 1080, 1024, (640, 300, 704, 364)),
    ("C4_upper", 7, 1920, 1080, 4096, (560, 790, 624, 854)),         # stream ids up to 2^32.6
    ("C4_spheres", 7, 1920, 1080, 4096, (1100, 640, 1164, 704)),
    ("C5_upper", 8, 3840, 2160, 4096, (1888, 2030, 1952, 2094)),      # stream ids up to 2^34.94 (the corners of the frame see nothing)
    ("C5_bunny", 8, 3840, 2160, 4096, (1480, 900, 1544, 964)),
    ("C1_rows", 0, 500, 500, 16, (0, 236, 500, 264)),
]


@pytest.mark.parametrize("name,scene,W,H,spp,crop", WINDOWS)
def test_true_size_window_parity(name, scene, W, H, spp, crop):
    g = np.load(os.path.join(GOLDEN, f"window_{name}.npz"))
    assert (int(g["width"]), int(g["height"]), int(g["spp"])) == (W, H, spp) and tuple(int(v) for v in g["crop"]) == crop
    hs = api.HostScene(scene, W, H)
    r = api.Renderer(hs, 0)
    try:
        r.render_async(W, H, spp, crop=crop)
        st = r.stats()
        acc = r.readback()
    finally:
        r.close(); hs.close()
    x0, y0, x1, y1 = crop
    assert acc.shape == (y1 - y0, x1 - x0, 4)
    assert st["paths"] == (x1 - x0) * (y1 - y0) * spp
    # identical streams: the trace() call count is equal up to rare ulp-flipped decisions, >= 99.9 % of the window's pixels
    # within 1e-4 (a window has few pixels but 1024-4096 samples each, so a single flipped path moves a pixel by < 1e-3)
    assert abs(int(st["rays"]) - int(g["rays"])) <= 2e-4 * int(g["rays"]), (st["rays"], int(g["rays"]))
    np.testing.assert_array_equal(acc[..., 3], g["acc"][..., 3])
    res = accfile.compare(accfile.finalize(acc), accfile.finalize(g["acc"]), rel=REL_TOL)
    assert res["frac_ok"] >= MIN_FRAC, res


@needs_ref
def test_crop_window_equals_full_frame():
    """A window is bit-identical to the same pixels of a full render (same streams, same per-pixel sum order)."""
    w, h, spp = 96, 54, 36
    full, _ = _gpu_render(7, w, h, spp)
    crop = (40, 10, 77, 31)
    part, st = _gpu_render(7, w, h, spp, crop=crop)
    np.testing.assert_array_equal(part, full[10:31, 40:77])
    ref, meta = oracle_util.ref_render(7, w, h, spp, crop=crop)
    assert st["rays"] == meta["rays"]
    with pytest.raises(api.MrtError):
        _gpu_render(7, w, h, spp, crop=(10, 10, 10, 20))


@pytest.mark.parametrize("scene,w,h,spp", [(0, 160, 160, 16), (1, 96, 96, 36), (7, 160, 90, 36), (8, 160, 90, 36)])
def test_cooperative_tree_traversal_is_bit_identical(scene, w, h, spp):
    """Mode B can traverse BVH trees warp-cooperatively (coop_tree.cuh: shared work stack, the hit of the leaf with the
    lowest depth-first rank wins; the default for big trees = the triangle meshes of scene 8).  Every box / primitive test
    is the per-lane traversal's, so the accumulator is bit-identical to MrtTuning.coop_trees = 1 (per-lane depth-first
    traversal) and the trace() count is equal; the leaf batch size does not matter either."""
    lane, st1 = _gpu_render(scene, w, h, spp, tuning=dict(coop_trees=1))
    coop, st2 = _gpu_render(scene, w, h, spp, tuning=dict(coop_trees=2))
    dflt, st0 = _gpu_render(scene, w, h, spp)
    assert (st1["coop_trees"], st2["coop_trees"], st0["coop_trees"]) == (0, 1, 1 if scene == 8 else 0)
    odd, st3 = _gpu_render(scene, w, h, spp, tuning=dict(coop_trees=2, coop_leaf_batch=5))
    np.testing.assert_array_equal(odd, lane)
    assert st2["coop_node_steps"] > 0 and st2["coop_leaf_steps"] > 0 and st1["coop_node_steps"] == 0
    assert st1["rays"] == st2["rays"] == st0["rays"]
    np.testing.assert_array_equal(coop, lane)
    np.testing.assert_array_equal(dflt, lane)


def test_cooperative_traversal_not_applicable():
    """Scenes without BVH trees (Cornell box) cannot ask for the cooperative traversal; the default silently does not use it."""
    _, st = _gpu_render(5, 64, 36, 16)
    assert st["coop_trees"] == 0
    with pytest.raises(api.MrtError):
        _gpu_render(5, 64, 36, 16, tuning=dict(coop_trees=2))


@pytest.mark.parametrize("scene", [5, 0, 7])
def test_tonemap_matches_reference(scene):
    """mrt_gpu_tonemap_device against the reference's own preview tone map (main.cpp:416-444 + ARGB32, vec3.h:327-333): the golden
    pair is the reference's final linear frame and the ARGB buffer ITS main loop computed from it (tests/golden/make_golden_tonemap.py).
    logf / log10f / powf differ by an ulp or two between the host libm and CUDA, so a channel may land on the other side of an
    integer boundary: +-1 LSB per channel, and at least 99 % of the channels exactly equal."""
    import torch
    g = np.load(os.path.join(GOLDEN, f"tonemap_scene{scene}.npz"))
    lin, want = g["linear"], g["argb"]
    h, w = want.shape
    img = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda:0")
    img[..., :3] = torch.from_numpy(lin).cuda()
    img[..., 3] = 1.0
    out = torch.zeros((h, w), dtype=torch.int32, device="cuda:0")
    hs = api.HostScene(5, w, h)
    r = api.Renderer(hs, 0)
    try:
        r.set_stream(torch.cuda.current_stream().cuda_stream)
        r.tonemap_device(img.data_ptr(), out.data_ptr(), w, h)
        torch.cuda.synchronize()
    finally:
        r.close(); hs.close()
    got = out.cpu().numpy().view(np.uint32)
    assert (got >> 24).max() == 0 and (want >> 24).max() == 0
    exact = 0
    for shift in (16, 8, 0):
        a, b = ((got >> shift) & 255).astype(np.int32), ((want >> shift) & 255).astype(np.int32)
        assert np.abs(a - b).max() <= 1, (shift, np.abs(a - b).max())
        exact += int((a == b).sum())
    assert exact >= 0.99 * 3 * h * w, exact / (3 * h * w)
    assert want.max() > 0


@pytest.mark.parametrize("devices", [(0, 0, 0), (0, 1)])
def test_reduce_finalize_over_peer_memory(devices):
    """mrt_gpu_reduce_finalize: n scenes that rendered disjoint sample slices of the same frame are summed in scene order,
    finalised (mean + luminance clamp) and tone-mapped on the GPUs.  (0, 0, 0): three scenes on one device (always runs);
    (0, 1): two GPUs, the stripes cross NVLink as peer loads / stores (needs gpurun --gpus 2)."""
    import torch
    if max(devices) >= torch.cuda.device_count():
        pytest.skip("needs %d GPUs" % (max(devices) + 1))
    w, h, spp = 160, 90, 36
    n = len(devices)
    hs = api.HostScene(7, w, h)
    rs = [api.Renderer(hs, d) for d in devices]
    try:
        parts = []
        for k, r in enumerate(rs):
            r.render_async(w, h, spp, sample_begin=spp * k // n, sample_end=spp * (k + 1) // n, max_luminance=2.0)
        for r in rs:
            parts.append(r.readback())
        img, argb = api.reduce_finalize(rs, max_luminance=2.0, tonemap=True)
        one = api.Renderer(hs, devices[0])
        one.render_async(w, h, spp, max_luminance=2.0)
        full = one.readback()
        argb_one = one.tonemap()
        one.close()
    finally:
        for r in rs:
            r.close()
        hs.close()
    acc = parts[0].copy()
    for p in parts[1:]:
        acc += p                                   # scene order, float32: what the kernel does
    np.testing.assert_array_equal(img[..., 3], full[..., 3])
    np.testing.assert_allclose(img[..., :3], accfile.finalize(acc, 2.0), rtol=1e-6, atol=1e-7)
    res = accfile.compare(img[..., :3], accfile.finalize(full, 2.0), rel=1e-5)
    assert res["n_bad"] == 0, res
    d = np.abs(((argb >> 8) & 255).astype(np.int32) - ((argb_one >> 8) & 255).astype(np.int32))
    assert d.max() <= 1


@needs_ref
@pytest.mark.parametrize("scene,w,h,spp,maxlum", [(5, 96, 54, 16, 0.5), (0, 64, 64, 9, 1000.0)])
def test_draw2_running_mean_matches_reference(scene, w, h, spp, maxlum):
    """mrt_gpu_render_running_mean = the reference's default mode (draw2, main.cpp:193-243): one-sample passes, running mean, the
    luminance clamp after every pass feeding back into the mean, a non-finite sample replaced by the mean so far.  Oracle:
    `mrt_ref render -draw2 1` (the same pixel update around the reference's trace(), identical streams)."""
    ref, _ = oracle_util.ref_render(scene, w, h, spp, draw2=True, maxlum=maxlum)
    hs = api.HostScene(scene, w, h)
    r = api.Renderer(hs, 0)
    try:
        mean = r.render_running_mean(w, h, spp, max_luminance=maxlum)
        st = r.stats()
        again = r.readback(finalize=True)           # the image buffer now IS the running mean
    finally:
        r.close(); hs.close()
    assert st["paths"] == w * h * spp and st["mode"] == api.MODE_BINNED
    np.testing.assert_array_equal(mean[..., 3], np.float32(spp))
    np.testing.assert_array_equal(again, mean)
    res = accfile.compare(mean[..., :3], ref[..., :3], rel=REL_TOL)
    assert res["frac_ok"] >= MIN_FRAC, res


def test_triangle_scene_object_golden():
    """triangle_scene_object (triangle.cpp:5-175, used by no stock scene): two lone triangles in the Cornell box
    (MRT_SCENE_EXTRA_TRIANGLES; oracle `-extra triangles`) against the reference; runs the everything-compiled-in kernel variant."""
    g = np.load(os.path.join(GOLDEN, "golden_extra_triangles_scene5.npz"))
    acc, st = _gpu_render(5 | api.SCENE_EXTRA_TRIANGLES, int(g["width"]), int(g["height"]), int(g["spp"]), int(g["depth"]))
    _check(acc, g["acc"], int(g["rays"]), st)
    assert int(st["rays"]) == int(g["rays"])
