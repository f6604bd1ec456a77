"""The bench JSON line contract, checked on the lines committed under profiles/ (CPU only): every key the
driver reads is there, with the meaning the contract gives it. A regression of bench.py's output shows
here before it costs a GPU run."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(path):
    txt = open(path).read().strip().splitlines()
    return json.loads([l for l in txt if l.startswith("{")][-1])


@pytest.mark.parametrize("name", ["r01_v20_bench_c2.json", "r01_v20_bench_c3.json"])
def test_own_arm_line(name):
    d = _line(os.path.join(ROOT, "profiles", name))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["metric"] == "ICP iterations/sec" and d["unit"] == "iterations/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"  # BASELINE.md publishes no number
    assert d["config"]["workload"] in ("C2", "C3") and "model" not in d["config"]
    assert d["warmup"] >= 3 and "flushed" in d["config"]["l2"]
    assert abs(d["value"] - d["n_gpus"] * 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]  # the host path cannot be faster than the resident one
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - r["algorithmic_bytes"] / (r["kernel_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["unit"] == d["unit"] and c["value"] > 0 and c["sample"]
    assert d["gpu_launches"] >= 1
    k = d["clocks"]
    assert k["sm_mhz"] > 0 and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line():
    d = _line(os.path.join(ROOT, "profiles", "r01_v20_bench_ref_c2.json"))
    assert d["impl"] == "reference" and d["metric"] == "ICP iterations/sec" and d["unit"] == "iterations/s"
    assert d["config"]["workload"] == "C2" and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_multi_gpu_lines_are_weak_scaling_aggregates():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_v20_bench_C2_n*_peer.json")))
    assert files
    for f in files:
        d = _line(f)
        assert d["n_gpus"] in (2, 4) and d["scaling"] == "weak" and d["config"]["collectives"].startswith("own kernels")
        assert abs(d["value"] - d["n_gpus"] * 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]  # whole-job aggregate over all ranks
