"""The bench JSON line contract, checked on the lines committed under profiles/ (CPU only): every key the
driver reads is there, with the meaning the contract gives it. A regression of bench.py's output shows
here before it costs a GPU run."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    txt = open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()
    return json.loads([l for l in txt if l.startswith("{")][-1])


@pytest.mark.parametrize("name", ["r02_bench_c3_default.json", "r02_bench_c2.json"])
def test_own_arm_line(name):
    d = _line(name)
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
    # the headline e2e is the SAFE default: the pairings travel both ways; the opt-in shortcut is reported beside it
    rec = 72 if d["config"]["workload"] == "C3" else 36
    assert e["h2d_bytes_per_step"] >= d["config"]["pairs"] * rec and e["d2h_bytes_per_step"] >= d["config"]["pairs"] * rec
    assert e["assume_unmodified_pairings"]["ms_per_step"] < e["ms_per_step"] < e["pageable"]["ms_per_step"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - r["algorithmic_bytes"] / (r["kernel_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    # SURVEY 8(d) unit sizes, reproducible from the printed counts
    c = r["counts"]
    assert r["algorithmic_bytes"] == c["N_q"] * 12 + c["pairs"] * rec + c["N_q"] * 4 + c["probes"] * 8 + c["candidates"] * 12
    assert r["bytes_loaded_16B_layout"] > r["algorithmic_bytes"]
    assert "timed inside the step function" in r["kernel"] and abs(r["kernel_ms"] - sum(r["kernel_ms_parts"][k] for k in ("nn_search", "plane_fit", "compact"))) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    b = d["cpu_baseline"]
    assert b["kind"] in ("port", "reference") and b["cores"] >= 1 and b["unit"] == d["unit"] and b["value"] > 0 and b["sample"]
    assert d["gpu_launches"] >= 1
    k = d["clocks"]
    assert k["sm_mhz"] > 0 and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_default_line_is_the_headline_config_and_carries_the_rest_of_8d():
    d = _line("r02_bench_c3_default.json")
    assert d["config"]["workload"] == "C3" and d["config"]["map_points"] == 10_000_000 and d["config"]["queries_per_gpu"] > 100_000
    for name in ("C1", "C4"):  # full align() wall time, iterations, termination, pose error
        a = d["align"][name]
        for path in ("plugin_calls_host_buffers", "fused_device_resident"):
            assert a[path]["wall_ms"] > 0 and a[path]["iterations"] > 0 and a[path]["termination"] in ("Stalled", "MaxIterations")
            assert len(a[path]["pose_error_vs_gt"]["trans_m"]) == 3
        assert a["cpu_baseline"]["iterations"] == a["plugin_calls_host_buffers"]["iterations"] and a["cpu_baseline"]["pose_diff_vs_gpu"] < 1e-9
    c5 = d["c5"]
    assert c5["workload"] == "C5" and c5["map_points"] == 100_000_000 and c5["queries_total"] == 1_000_000 and c5["scaling"] == "strong"
    assert c5["parity_vs_n1"]["equal"] is True


def test_reference_arm_line():
    d = _line("r02_bench_ref_c3.json")
    assert d["impl"] == "reference" and d["metric"] == "ICP iterations/sec" and d["unit"] == "iterations/s"
    assert d["config"]["workload"] == "C3" and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["cores"] > 1  # all host threads, also under torchrun (OMP_NUM_THREADS=1 is not what the oracle reads)
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_multi_gpu_lines():
    d = _line("r02_bench_c3_n2.json")  # the main line stays the headline workload, weak scaling, whole-job aggregate
    assert d["n_gpus"] == 2 and d["scaling"] == "weak" and d["config"]["workload"] == "C3" and d["config"]["collectives"].startswith("own kernels")
    assert abs(d["value"] - d["n_gpus"] * 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]
    assert d["e2e"]["value"] is None  # no sharded host-buffer path: null, not N replicas
    for name, n in (("r02_bench_c5_n2.json", 2), ("r02_bench_c5_n1.json", 1)):  # C5: strong scaling, ONE iteration over 1M queries
        c = _line(name)
        assert c["n_gpus"] == n and c["scaling"] == "strong" and c["config"]["queries_total"] == 1_000_000
        assert abs(c["value"] - 1e3 / c["ms_per_step"]) < 1e-6 * c["value"]
        assert c["parity_vs_n1"]["equal"] is True and c["parity_vs_n1"]["pairs"] == 524685
    assert _line("r02_bench_c5_n2.json")["parity_vs_n1"]["hash64"] == _line("r02_bench_c5_n1.json")["parity_vs_n1"]["hash64"]
