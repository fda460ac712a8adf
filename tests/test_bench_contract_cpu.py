"""The bench.py JSON contract, checked on the lines committed under profiles/ (produced on a B200 by the final build of the round):
a guard against dropping or renaming a key the driver and the judge read."""
import json
import os

import pytest

PROF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")


def _line(name):
    path = os.path.join(PROF, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not committed yet")
    return json.loads(open(path).read().strip().splitlines()[-1])


def test_ours_line_has_the_contract_keys():
    d = _line("r01_bench_ours_final.json")
    for k in ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"]:
        assert k in d, k
    assert d["unit"] == "frames/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] >= d["steps"] * 10
    assert abs(d["value"] - 1000.0 / d["ms_per_step"]) / d["value"] < 1e-6
    e = d["e2e"]
    assert e["unit"] == "frames/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] < d["value"] * 1.001            # host buffers + copies + a sync per frame cannot beat the device-resident loop
    r = d["roofline"]
    for k in ["bound", "achieved", "peak", "unit", "frac", "traffic"]:
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_line_has_the_contract_keys():
    d = _line("r01_bench_reference_final.json")
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["unit"] == "frames/s"
    assert "workload" in d["config"]
