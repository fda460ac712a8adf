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


@pytest.mark.parametrize("name", ["r01_bench_ours_final.json", "r02_bench_ours_final.json", "r02_bench_ours_cfg1.json",
                                  "r02_bench_ours_cfg2.json", "r02_bench_ours_cfg3.json"])
def test_ours_line_has_the_contract_keys(name):
    d = _line(name)
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


@pytest.mark.parametrize("name", ["r01_bench_reference_final.json", "r02_bench_reference_final.json", "r02_bench_reference_cfg1.json",
                                  "r02_bench_reference_cfg2.json", "r02_bench_reference_cfg3.json", "r02_bench_reference_cfg5.json"])
def test_reference_line_has_the_contract_keys(name):
    d = _line(name)
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["unit"] == "frames/s"
    assert "workload" in d["config"]


@pytest.mark.parametrize("tag", ["final", "cfg1", "cfg2", "cfg3"])
def test_both_arms_time_the_same_workload(tag):
    """The two arms of a configuration are compared by the driver: they must describe the same workload, up to the number of
    points per frame (a mean over the frames each arm timed: 50 after 10 warm-up frames against 50 after 5)."""
    import re
    ours, ref = _line(f"r02_bench_ours_{tag}.json"), _line(f"r02_bench_reference_{tag}.json")
    strip = lambda w: re.sub(r"\d+ points/frame", "N points/frame", w)
    assert strip(ours["config"]["workload"]) == strip(ref["config"]["workload"])
    assert ours["metric"] == ref["metric"] and ours["unit"] == ref["unit"] and ours["higher_is_better"] == ref["higher_is_better"]


@pytest.mark.parametrize("n", [2, 4, 8])
def test_sharded_lines(n):
    """bench.py --gpus N times ONE volume sharded over N GPUs (strong scaling) and carries rank 0's single-GPU time for the same
    frames; the reference arm of the same launch runs the same cfg5 volume."""
    d = _line(f"r02_bench_sharded_n{n}.json")
    assert d["n_gpus"] == n and d["scaling"] == "strong" and d["unit"] == "frames/s" and d["gpu_launches"] > 0
    assert abs(d["value"] - 1000.0 / d["ms_per_step"]) / d["value"] < 1e-6
    single = d["single_gpu_same_workload"]
    assert abs(d["strong_scaling"]["speedup"] - single["ms_per_step"] / d["ms_per_step"]) < 1e-6
    assert abs(d["strong_scaling"]["efficiency"] - d["strong_scaling"]["speedup"] / n) < 1e-9
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and 0 < d["e2e"]["value"] <= d["value"] * 1.001
    assert d["broadcast_bytes_per_frame"] > 0 and d["slab_sweeps_ms_max_over_ranks"] < d["ms_per_step"]
    ref = _line("r02_bench_reference_cfg5.json")
    assert ref["config"]["workload"].split(",")[0] == d["config"]["workload"].split(",")[0]      # "cfg5: 1024x1024x1016 @ 0.1 m"
