"""The command-line front end (the reference's main() with the GPU renderer behind the C ABI)."""
import os
import subprocess

import numpy as np
import pytest

import oracle_util
from miniraytracer_b200 import accfile, build

EXE = build.EXE


def test_help_lists_reference_options():
    out = subprocess.run([EXE, "--help"], capture_output=True, text=True, check=True).stdout
    for opt in ("-width", "-height", "-samples", "-depth", "-maxlum", "-threads", "-tilesize", "-mode", "-scene", "-delay"):
        assert opt in out     # cmdline_parser.cpp:107-122
    for opt in ("-gpus", "-seed", "-assets", "-out"):
        assert opt in out


def _read_pfm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"PF"
        w, h = (int(v) for v in f.readline().split())
        assert float(f.readline()) < 0
        return np.fromfile(f, dtype="<f4", count=w * h * 3).reshape(h, w, 3)


@pytest.mark.gpu
@pytest.mark.skipif(not oracle_util.have_ref(), reason="oracle/_ref/mrt_ref not on this box")
@pytest.mark.parametrize("mode", [0, 1])
def test_cli_render_matches_oracle(tmp_path, mode):
    w, h, spp = 160, 90, 64
    out = tmp_path / "img.pfm"
    r = subprocess.run([EXE, "-scene", "5", "-width", str(w), "-height", str(h), "-samples", str(spp), "-mode", str(mode),
                        "-assets", oracle_util.ASSETS, "-out", str(out)], capture_output=True, text=True, check=True)
    assert "Mrays/s" in r.stdout
    img = _read_pfm(out)
    ref = accfile.finalize(oracle_util.ref_render(5, w, h, spp)[0])
    res = accfile.compare(img, ref, rel=1e-4)
    assert res["frac_ok"] >= 0.999, res
    ppm = tmp_path / "img.ppm"
    subprocess.run([EXE, "-scene", "5", "-width", "64", "-height", "36", "-samples", "4", "-assets", oracle_util.ASSETS, "-out", str(ppm)],
                   capture_output=True, text=True, check=True)
    assert ppm.read_bytes().startswith(b"P6\n64 36\n255\n")


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(not oracle_util.have_ref(), reason="oracle/_ref/mrt_ref not on this box")
def test_cli_two_gpus_reduce_over_nvlink(tmp_path):
    """mrt_b200 -gpus 2: each GPU renders half of the samples, the accumulators are summed + finalised + tone-mapped by
    mrt_gpu_reduce_finalize over peer memory (no host-side sum); the image matches the oracle like a one-GPU render."""
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    w, h, spp = 160, 90, 64
    out = tmp_path / "img2.pfm"
    subprocess.run([EXE, "-scene", "7", "-width", str(w), "-height", str(h), "-samples", str(spp), "-mode", "0", "-gpus", "2",
                    "-assets", oracle_util.ASSETS, "-out", str(out)], capture_output=True, text=True, check=True)
    img = _read_pfm(out)
    ref = accfile.finalize(oracle_util.ref_render(7, w, h, spp)[0])
    res = accfile.compare(img, ref, rel=1e-4)
    assert res["frac_ok"] >= 0.999, res
    ppm = tmp_path / "img2.ppm"
    subprocess.run([EXE, "-scene", "5", "-width", "64", "-height", "36", "-samples", "16", "-gpus", "2", "-assets", oracle_util.ASSETS,
                    "-out", str(ppm)], capture_output=True, text=True, check=True)
    one = tmp_path / "img1.ppm"
    subprocess.run([EXE, "-scene", "5", "-width", "64", "-height", "36", "-samples", "16", "-gpus", "1", "-assets", oracle_util.ASSETS,
                    "-out", str(one)], capture_output=True, text=True, check=True)
    a = np.frombuffer(ppm.read_bytes()[len(b"P6\n64 36\n255\n"):], dtype=np.uint8).astype(np.int32)
    b = np.frombuffer(one.read_bytes()[len(b"P6\n64 36\n255\n"):], dtype=np.uint8).astype(np.int32)
    assert np.abs(a - b).max() <= 1     # same image up to the float summation order of the two halves
