"""The bench lines committed under profiles/ carry every key the measurement contract names (they were printed by
bench.py on the GPU box; this keeps bench.py's output shape from drifting unnoticed)."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROFILES = os.path.join(ROOT, "profiles")


def line(name):
    with open(os.path.join(PROFILES, name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


@pytest.mark.parametrize("name,n", [("r1_bench_1gpu.json", 1), ("r1_bench_2gpu.json", 2), ("r1_bench_4gpu.json", 4),
                                    ("r1_bench_8gpu.json", 8)])
def test_bench_line_has_the_contract_keys(name, n):
    d = line(name)
    assert d["metric"] == "actor-updates/sec" and d["unit"] == "actor-updates/s" and d["n_gpus"] == n
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["steps"] >= 1 and d["warmup"] >= 3 and d["ms_per_step"] > 0 and d["value"] > 1e9
    assert d["gpu_launches"] >= 4 * d["steps"]                      # four kernels per tick at least
    e = d["e2e"]
    assert e["unit"] == d["unit"] and 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] == 0
    assert e["d2h_bytes_per_step"] >= 8 * 0.5 * d["config"]["actors"]          # a float pair per living actor
    c = d["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0 < r["frac"] < 1 and r["traffic"] is None or r["traffic"] > 0
    if n == 1:
        b = d["cpu_baseline"]
        assert b["kind"] == "port" and b["cores"] == 1 and b["value"] > 0 and b["sample"]
        assert abs(d["value"] - d["roofline"]["living_actors"] * d["steps"] / (d["ms_per_step"] * d["steps"] * 1e-3)) < 1e-3 * d["value"]


def test_reference_arm_line():
    d = line("r1_bench_reference.json")
    assert d["impl"] == "reference" and d["metric"] == "actor-updates/sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == line("r1_bench_1gpu.json")["config"]["workload"]


@pytest.mark.parametrize("name,n", [("r2/r2n_bench_1gpu.json", 1), ("r2/r2p_bench_2gpu_weak.json", 2), ("r2/r2m_bench_8gpu_weak.json", 8)])
def test_round_2_bench_line(name, n):
    """The final lines of round 2: the contract's keys, the delta-frame e2e beside the full-frame one, the oracle's digest."""
    d = line(name)
    assert d["metric"] == "actor-updates/sec" and d["unit"] == "actor-updates/s" and d["n_gpus"] == n
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["steps"] >= 1 and d["warmup"] >= 3 and d["value"] > 1e9 and d["gpu_launches"] >= 4 * d["steps"]
    e, f = d["e2e"], d["e2e_full_frames"]
    assert e["unit"] == d["unit"] and 0 < f["value"] < e["value"] < d["value"] and e["h2d_bytes_per_step"] == 0
    assert 0 < e["d2h_bytes_per_step"] < f["d2h_bytes_per_step"]               # a delta frame is smaller than a full one
    assert f["d2h_bytes_per_step"] >= 8 * 0.5 * d["config"]["actors"]          # a float pair per slot
    assert e["d2h_bytes_per_step"] >= 8 * e["positions_changed_per_step"]      # every changed position crossed PCIe
    r = d["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1 and r["traffic"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if n == 1:
        assert d["state_digest"]["matches_oracle"] is True
        assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] > 0


def test_round_2_reference_arm_names_what_it_ran():
    d = line("r2/r2n_bench_reference.json")
    assert d["impl"] == "reference" and d["cpu_baseline"]["value"] == d["value"] > 0
    assert (d["config"]["grid"], d["config"]["actors"]) == (256, 1000000)        # the sample, not the GPU arm's world
    assert d["config"]["sample_of"] == {"grid": 1024, "actors": 16000000}
