"""Pins the oracle (the reference's own code, patched only to build headless) against the only
known-answer data that exists for this path: the published PCG32 demo vector, the survey's probe
vectors and the committed golden renders (tests/golden, generated from the
reference by make_golden.py).  The reference ships no tests of its own."""
import os

import numpy as np
import pytest

import oracle_util
from miniraytracer_b200 import accfile

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _kat_lines():
    return open(os.path.join(GOLDEN, "kat.txt")).read().splitlines()


def test_pcg32_published_vector():
    # pcg32-demo, seed (42, 54): the first six outputs of the reference PCG32 implementation
    line = [l for l in _kat_lines() if l.startswith("pcg32 seed=42 seq=54")][0]
    assert line.split(":")[1].split() == ["a15c02b7", "7b47f409", "ba1d3330", "83d2f293", "bfa4784b", "cbed606e"]


def test_pcg32_main_seed_vectors():
    lines = _kat_lines()
    r32 = [l for l in lines if l.startswith("pcg32 mainseed rand32")][0].split(":")[1].split()
    assert r32 == ["5ee7b849", "6347ef22", "0ca09808", "5d6d4a33"]
    rf = [l for l in lines if l.startswith("pcg32 mainseed randf")][0].split(":")[1].split()
    vals = np.array([int(x, 16) for x in rf], dtype=np.uint32).view(np.float32)
    np.testing.assert_allclose(vals, [0.810311437, 0.561985254, 0.254639626, 0.85382688], rtol=0, atol=1e-9)


def test_left_to_right_draw_order():
    # random_in_sphere((42,54)) must consume the draws in x, y, z order (the reference's compiler order)
    lines = _kat_lines()
    sph = [l for l in lines if l.startswith("random_in_sphere")][0].split(":")[1].split()
    first = [l for l in lines if l.startswith("pcg32 seed=42 seq=54")][0].split(":")[1].split()
    r = np.array([int(x, 16) for x in first[:3]], dtype=np.uint32)
    f = ((r & 0x7FFFFF) | 0x3F800000).view(np.float32) - np.float32(1)
    p = np.float32(2) * f - np.float32(1)
    if float(np.dot(p, p)) < 1.0:   # first triple accepted
        got = np.array([int(x, 16) for x in sph], dtype=np.uint32).view(np.float32)
        np.testing.assert_array_equal(got, p)


@pytest.mark.skipif(not oracle_util.have_ref(), reason="oracle/_ref/mrt_ref not built")
def test_oracle_binary_matches_kat():
    assert oracle_util.ref_run(["kat"]).stdout == open(os.path.join(GOLDEN, "kat.txt")).read()


@pytest.mark.skipif(not oracle_util.have_ref(), reason="oracle/_ref/mrt_ref not built")
@pytest.mark.parametrize("scene", range(9))
def test_oracle_reproduces_golden(scene):
    g = np.load(os.path.join(GOLDEN, f"golden_scene{scene}.npz"))
    acc, meta = oracle_util.ref_render(scene, int(g["width"]), int(g["height"]), int(g["spp"]), int(g["depth"]))
    assert meta["rays"] == int(g["rays"])
    if scene in (0, 1, 2, 3, 4, 5, 6, 7, 8):
        # libm may differ by an ulp between hosts; everything else is integer/IEEE exact
        res = accfile.compare(accfile.finalize(acc), accfile.finalize(g["acc"]), rel=1e-5)
        assert res["frac_ok"] >= 0.999, res


@pytest.mark.skipif(not oracle_util.have_ref(), reason="oracle/_ref/mrt_ref not built")
def test_libm_canonicalisation_is_a_small_change():
    """The oracle binds sinf/cosf/atan2f/asinf/logf/powf to correctly rounded versions (oracle/cr_libm.cpp).
    With MRT_ORACLE_LIBM=host it uses the host libm's own float functions, i.e. the reference exactly as it
    would run here.  The two renders must be the same image up to the rare 1-ulp disagreements of libm:
    nearly all pixels bit-identical, the rest statistically equivalent."""
    import subprocess, tempfile
    outs = []
    for mode in ("cr", "host"):
        env = dict(os.environ)
        env["MRT_ORACLE_LIBM"] = mode
        path = os.path.join(tempfile.gettempdir(), f"mrt_libm_{mode}_{os.getpid()}.bin")
        subprocess.run([oracle_util.REF_BIN, "render", "-scene", "0", "-width", "120", "-height", "120", "-samples", "16",
                        "-out", path], cwd=oracle_util.RUN_DIR, env=env, check=True, capture_output=True)
        outs.append(accfile.read_acc(path)[0])
        os.unlink(path)
    cr, host = outs
    identical = float(np.all(cr == host, axis=-1).mean())
    assert identical > 0.95, identical
    res = accfile.compare(accfile.finalize(cr), accfile.finalize(host), rel=1e-4)
    assert res["frac_ok"] > 0.995, res


@pytest.mark.skipif(not oracle_util.have_ref(), reason="oracle/_ref/mrt_ref not built")
def test_draw2_running_mean_semantics():
    """`mrt_ref render -draw2 1` = the reference's default worker's pixel update (draw2, main.cpp:214-231): running mean in sample
    order, luminance clamp after EVERY sample feeding back into the mean.  Cross-check against a float32 numpy restatement fed
    with the per-sample radiance of the same streams (one oracle render per sample), with a clamp low enough to fire often."""
    from miniraytracer_b200 import accfile
    scene, w, h, spp, maxlum = 5, 24, 13, 9, 0.15
    got, _ = oracle_util.ref_render(scene, w, h, spp, draw2=True, maxlum=maxlum)
    mean = np.zeros((h, w, 3), dtype=np.float32)
    c709 = np.array([0.212655, 0.715158, 0.072187], dtype=np.float32)
    clamped = 0
    for s in range(spp):
        one, _ = oracle_util.ref_render(scene, w, h, spp, s0=s, s1=s + 1)
        finite = one[..., 3:4] > 0
        c = np.where(finite, one[..., :3], mean if s else np.float32(0)).astype(np.float32)
        if s:
            c = (mean + (c - mean) * (np.float32(1.0) / np.float32(s + 1.0))).astype(np.float32)
        prod = c * c709
        lum = (prod[..., 0] + prod[..., 1]) + prod[..., 2]
        over = lum > np.float32(maxlum)
        clamped += int(over.sum())
        with np.errstate(divide="ignore", invalid="ignore"):
            k = (np.float32(maxlum) / lum).astype(np.float32)
        mean = np.where(over[..., None], c * k[..., None], c).astype(np.float32)
    assert clamped > 10
    np.testing.assert_array_equal(got[..., :3], mean)
    plain = accfile.finalize(oracle_util.ref_render(scene, w, h, spp)[0], maxlum)
    assert np.abs(plain - mean).max() > 1e-3      # the per-pass clamp is NOT the same as clamping the final mean
