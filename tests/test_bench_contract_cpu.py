"""The committed bench lines (profiles/r1_bench_*.json, produced by bench.py on a B200) carry every key of the
measurement contract; bench.py's workload table is consistent with SURVEY.md 8(d)."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = sorted(f for f in glob.glob(os.path.join(ROOT, "profiles", "r1_bench_*.json")) if "reference" not in f and "_n" not in os.path.basename(f))


@pytest.mark.parametrize("path", OURS, ids=[os.path.basename(p) for p in OURS])
def test_bench_line_contract(path):
    d = json.load(open(path))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["unit"] == "gradient-steps/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"] and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.05
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert abs(d["value"] * d["ms_per_step"] / 1000.0 - d["n_gpus"]) < 0.02 * d["n_gpus"]      # value == n_gpus / step time


def test_reference_arm_line():
    d = json.load(open(os.path.join(ROOT, "profiles", "r1_bench_reference_sac_hopper.json")))
    assert d["impl"] == "reference" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]


def test_algorithmic_bytes_match_the_survey():
    import bench

    exp = {"sac_hopper": 6_190_312, "sac_ant": 8_796_632, "gail_walker": 6_947_216, "td3_humanoid": 15_171_912}
    for name, want in exp.items():
        assert int(bench.algorithmic_bytes_per_step(bench.WORKLOADS[name])) == want, name


def test_round2_default_line_covers_the_other_baseline_configs():
    """the default bench line carries GAIL Walker / TD3 Humanoid / SAC Ant sub-records measured the same way,
    the real launch size, and (N > 1) the replica check."""
    for fn in ("r2_bench_sac_hopper.json", "r2_bench_sac_hopper_steps20.json"):
        d = json.load(open(os.path.join(ROOT, "profiles", fn)))
        assert set(d["workloads"]) == {"gail_walker", "td3_humanoid", "sac_ant"}
        for name, r in d["workloads"].items():
            for k in ("value", "ms_per_step", "e2e", "roofline", "cpu_baseline", "gpu_launches"):
                assert k in r, (name, k)
            assert r["e2e"]["value"] > 0 and r["gpu_launches"] > 0
        launch = min(1000, d["steps"])
        assert d["roofline"]["launch_steps"] == launch and ("%d gradient steps per kernel launch" % launch) in d["config"]["workload"]
        assert d["cpu_baseline"]["kind"] == "reference"          # the reference's own rlkit classes (baseline/_ref)
        assert d["clocks"]["samples"] > 0
    for fn, n in (("r2_bench_sac_hopper_n2.json", 2), ("r2_bench_sac_hopper_n8.json", 8)):
        d = json.load(open(os.path.join(ROOT, "profiles", fn)))
        assert d["n_gpus"] == n and set(d["workloads"]) == {"sac_ant"}
        for name, c in d["replica_check"].items():
            assert c["equal"] and c["ok"] and c["world"] == n, (fn, name, c)
