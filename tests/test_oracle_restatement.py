"""The plain-C restatement (oracle/mrt_oracle.c) against the reference itself (oracle/_ref/mrt_ref): it reads
the scene dump the REFERENCE prints and must reproduce the reference's accumulator exactly -- same trace()
count, same finite-sample counts, bit-identical radiance sums (both are IEEE float32 in the same operation
order, built with -ffp-contract=off, same correctly rounded libm)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oracle_util
from miniraytracer_b200 import accfile

ROOT = oracle_util.ROOT
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
REST = os.path.join(ROOT, "oracle", "_ref", "mrt_oracle")


@pytest.fixture(scope="module")
def restatement():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "restatement"], check=True, capture_output=True)
    return REST


def _run(exe, dump, w, h, spp, sky, depth=32, extra=()):
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        out = f.name
    try:
        subprocess.run([exe, "-dump", dump, "-image", os.path.join(oracle_util.ASSETS, "earthmap.ppm"), "-width", str(w), "-height", str(h),
                        "-samples", str(spp), "-depth", str(depth), "-sky", str(sky), "-out", out, *extra], check=True, capture_output=True)
        return accfile.read_acc(out)
    finally:
        os.unlink(out)


@pytest.mark.skipif(not oracle_util.have_ref(), reason="oracle/_ref/mrt_ref not built")
@pytest.mark.parametrize("scene", range(9))
def test_restatement_bit_identical_to_reference(restatement, scene, tmp_path):
    w, h, spp = 96, 64, 9
    dump = str(tmp_path / "scene.txt")
    oracle_util.ref_dump_scene(scene, w, h, dump)
    ref, rmeta = oracle_util.ref_render(scene, w, h, spp)
    acc, meta = _run(restatement, dump, w, h, spp, 1 if scene < 5 else 0)
    assert meta["rays"] == rmeta["rays"]
    np.testing.assert_array_equal(acc, ref)


@pytest.mark.skipif(not oracle_util.have_ref(), reason="oracle/_ref/mrt_ref not built")
def test_restatement_sample_slices_and_depth(restatement, tmp_path):
    dump = str(tmp_path / "scene.txt")
    oracle_util.ref_dump_scene(5, 64, 36, dump)
    ref, _ = oracle_util.ref_render(5, 64, 36, 16, depth=4, s0=4, s1=12)
    acc, _ = _run(restatement, dump, 64, 36, 16, 0, depth=4, extra=("-s0", "4", "-s1", "12"))
    np.testing.assert_array_equal(acc, ref)


@pytest.mark.skipif(not oracle_util.have_ref(), reason="oracle/_ref/mrt_ref not built")
def test_restatement_crop_window_with_large_stream_ids(restatement, tmp_path):
    """Crop window of C5's true frame (3840x2160, 4096 spp): stream ids (y*W+x)*N+s near 2^35."""
    dump = str(tmp_path / "scene.txt")
    oracle_util.ref_dump_scene(8, 3840, 2160, dump)
    crop = (1900, 2040, 1906, 2042)
    ref, rmeta = oracle_util.ref_render(8, 3840, 2160, 4096, s0=4000, s1=4024, crop=crop)
    acc, meta = _run(restatement, dump, 3840, 2160, 4096, 0, extra=("-s0", "4000", "-s1", "4024", "-x0", "1900", "-y0", "2040", "-x1", "1906", "-y1", "2042"))
    assert meta["rays"] == rmeta["rays"]
    np.testing.assert_array_equal(acc, ref)


@pytest.mark.skipif(not oracle_util.have_ref(), reason="oracle/_ref/mrt_ref not built")
@pytest.mark.parametrize("scene", [5, 7])
def test_restatement_sphere_light(restatement, scene, tmp_path):
    """Light list with both allocated entries (ceiling light + glass sphere): sphere pdf_value / pdf_generate, random_towards_sphere."""
    dump = str(tmp_path / "scene.txt")
    oracle_util.ref_dump_scene(scene, 96, 54, dump, all_lights=True)
    ref, rmeta = oracle_util.ref_render(scene, 96, 54, 9, all_lights=True)
    acc, meta = _run(restatement, dump, 96, 54, 9, 0)
    assert meta["rays"] == rmeta["rays"]
    np.testing.assert_array_equal(acc, ref)


@pytest.mark.parametrize("scene", [2, 3, 5, 6])
def test_restatement_on_committed_dumps(restatement, scene):
    """Runs without the reference binary: committed scene dumps (printed by the reference) + committed golden
    renders (rendered by the reference) at 64x36x4."""
    # the committed dumps were printed for a 200x160 frame; the camera depends on the aspect only through
    # horz/llcorner, so render that frame size and compare with a golden of the same size generated alongside
    g = np.load(os.path.join(GOLDEN, f"golden_restatement_scene{scene}.npz"))
    acc, meta = _run(restatement, os.path.join(GOLDEN, f"scene{scene}_dump_200x160.txt"), 200, 160, 4, 1 if scene < 5 else 0)
    assert meta["rays"] == int(g["rays"])
    res = accfile.compare(accfile.finalize(acc), accfile.finalize(g["acc"]), rel=1e-5)
    assert res["n_bad"] == 0, res
