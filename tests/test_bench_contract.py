"""The committed bench lines (profiles/r02_bench_*.json: what `python bench.py` printed on a B200) carry every key the measurement
contract names, with consistent values. Guards the evidence files and bench.py's output format against drifting apart; needs no GPU."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASELINE = json.load(open(os.path.join(ROOT, "BASELINE.json")))


def _line(name):
    return json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])


@pytest.mark.parametrize("name", ["r02_bench_n1.json", "r02_bench_n1_hdri-test.json", "r02_scale_n8_cornell-lucy.json"])
def test_bench_line_has_the_contract_keys(name):
    d = _line(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert k in d, k
    assert d["metric"] == d["unit"] == "Mpaths/s" and d["higher_is_better"] is True and d["scaling"] == "strong" and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["steps"] >= 1 and d["value"] > 0 and d["gpu_launches"] > 0 and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["dtype"] == "f64 geometry / f32 radiance"
    e = d["e2e"]
    assert e["unit"] == d["unit"] and 0 < e["value"] <= d["value"] * 1.001 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    if d["n_gpus"] == 1 and name == "r02_bench_n1.json":
        cpu = d["cpu_baseline"]
        for k in ("value", "unit", "cores", "kind", "sample"):
            assert k in cpu, k
        assert cpu["kind"] in ("port", "reference") and cpu["unit"] == d["unit"] and cpu["cores"] >= 1
        assert d["config"]["workload"] == "cornell-lucy"    # the configuration BASELINE.json quotes its metric on


def test_reference_arm_line():
    d = _line("r02_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == d["unit"] == "Mpaths/s" and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("port", "reference")
    mine = _line("r02_bench_n1.json")
    assert d["config"]["workload"] == mine["config"]["workload"] and d["higher_is_better"] == mine["higher_is_better"]


def test_baseline_json_names_the_benched_metric():
    text = json.dumps(BASELINE).lower()
    assert "paths" in text and "cornell-lucy" in text
