"""The JSON lines bench.py printed on the B200 (committed under profiles/) carry every key the driver's contract names, and their
derived fields are consistent with each other. Guards the contract on CPU: a change of bench.py that drops a key shows up when
the evidence is refreshed."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROFILES = os.path.join(ROOT, "profiles")


def _line(name):
    path = os.path.join(PROFILES, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not captured yet")
    return json.loads(open(path).read().strip().splitlines()[-1])


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_headline_line_has_the_contract_keys(n):
    d = _line(f"r2b_bench_headline_n{n}.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "parity", "build", "kitchen"):
        assert k in d, k
    assert d["n_gpus"] == n and d["scaling"] == "strong" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["unit"] == "Mrays/s" and "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "frac_algorithmic", "frac_dram", "traffic"):
        assert k in r, k
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["unit"] == "GB/s"
    assert r["bound"] != "hbm"  # ncu: the kernel is issue / pipe bound, its DRAM share is ~10 %
    if n == 1:
        assert r["traffic"] and abs(r["frac_dram"] - r["traffic"] / (r["launch_ms"] * 1e-3) / 1e9 / r["peak"]) < 1e-6
        c = d["cpu_baseline"]
        assert c and c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = d["e2e"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in e, k
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    p = d["parity"]
    assert d["parity_ok"] is True and p["replicas_identical_across_ranks"] and p["probe_hits_identical_across_ranks"]
    assert d["clocks"]["sm_mhz"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    k = d["kitchen"]
    assert k["roofline"]["bound"] == "issue/L2" and k["value"] > 0 and k["e2e"]["value"] > 0


def test_reference_arm_line():
    d = _line("r2b_bench_headline_reference_arm.json")
    ours = _line("r2b_bench_headline_n1.json")
    assert d["impl"] == "reference" and d["metric"] == ours["metric"] and d["unit"] == ours["unit"] and d["config"] == ours["config"]
    assert d["higher_is_better"] == ours["higher_is_better"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_traffic_file_matches_the_headline_launch():
    t = json.load(open(os.path.join(PROFILES, "traffic.json")))
    d = _line("r2b_bench_headline_n1.json")
    # (the bench line is printed before the ncu capture of the same evidence run, so it carries the previous capture's bytes)
    assert t["s3"]["rays"] == d["config"]["rays_total"] and abs(t["s3"]["bytes"] - d["roofline"]["traffic"]) < 0.01 * t["s3"]["bytes"]
    assert os.path.exists(os.path.join(ROOT, t["s3"]["source"])) and os.path.exists(os.path.join(ROOT, t["kitchen"]["source"]))
