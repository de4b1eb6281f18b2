"""bench.py's contract with the driver (one JSON line, stated keys) -- the reference arm on the CPU here; our arm on a GPU.
The reference arm times the reference's OWN renderer (oracle/_ref, built from /root/reference by oracle/build_ref.sh); our arm
has no CPU path and must say so."""
import json
import os
import subprocess
import sys

import pytest

import oracle_util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")
needs_ref = pytest.mark.skipif(not oracle_util.have_ref(), reason="oracle/_ref/mrt_ref not built")

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
             "data", "config", "cpu_baseline", "e2e"}


def _line(stdout):
    lines = [l for l in stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, stdout
    return json.loads(lines[0])


@needs_ref
def test_reference_arm_line_and_rank_filter():
    env = dict(os.environ, RANK="0")
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "1", "--warmup", "0", "--no-per-config", "--cpu-spp", "1"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr
    d = _line(r.stdout)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "Mpath-samples/s" and d["unit"] == "Mpaths/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["config"]["width"] == 1920 and d["config"]["height"] == 1080 and d["config"]["spp"] == 1024 and d["config"]["max_bounces"] == 32
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # under torchrun only rank 0 works and prints
    r1 = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                        env=dict(os.environ, RANK="1", WORLD_SIZE="2"), timeout=120)
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_both_arms_describe_the_same_config():
    sys.path.insert(0, ROOT)
    import bench
    assert set(bench.WORKLOADS) == {"C1", "C2", "C3", "C4", "C5"}          # BASELINE.json configs[0..4]
    for name, (scene, w, h, spp, depth) in bench.WORKLOADS.items():
        c = bench.config_dict(name, scene, w, h, spp, depth)
        assert set(c) == {"workload", "scene", "width", "height", "spp", "max_bounces"} and c["workload"].startswith(name + ":")
    baseline = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert bench.WORKLOADS["C2"][1:4] == (1920, 1080, 1024), baseline.get("metric")


def test_our_arm_has_no_cpu_path():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, BENCH, "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_our_arm_line_on_the_gpu():
    r = subprocess.run([sys.executable, BENCH, "--steps", "2", "--warmup", "3", "--no-per-config", "--no-cpu-baseline", "--spp", "64"],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, MRT_NO_BUILD="1"))
    assert r.returncode == 0, r.stderr
    d = _line(r.stdout)
    assert (BASE_KEYS - {"cpu_baseline"}) <= set(d) and "impl" not in d
    assert d["value"] > 0 and d["gpu_launches"] == 4 and d["reduced"]                     # --spp marks the line as reduced
    roof = d["roofline"]
    assert 0 < roof["frac"] < 1 and roof["achieved"] / roof["peak"] == pytest.approx(roof["frac"]) and roof["unit"] == "TFLOP/s"
    e = d["e2e"]
    assert e["value"] > 0 and e["d2h_bytes_per_step"] == 1920 * 1080 * 16 and e["h2d_bytes_per_step"] > 0
    assert 0 < d["warp_busy_frac"] <= 1 and 0 < d["stage_sum_frac"] < 1
    assert "sm_mhz" in d["clocks"] and "reasons" in d["clocks"]
