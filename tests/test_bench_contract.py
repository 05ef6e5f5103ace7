"""bench.py's output contract, checked on the committed lines of the last GPU run (profiles/r02_bench_*.json): every
key the driver and the judge read is present and consistent.  (The lines themselves are produced on the GPU box.)"""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not committed")
    return json.load(open(path))


def test_ours_line_has_the_contract_keys():
    b = load("r02_bench_n1.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in b, k
    assert b["impl"] == "ours" and b["n_gpus"] == 1 and b["warmup"] >= 3 and b["higher_is_better"] is True
    assert b["unit"] == "Mrays/s" and b["dtype"] == "f32" and b["data"] == "synthetic" and b["vs_baseline"] is None
    assert "workload" in b["config"] and "model" not in b["config"]
    assert abs(b["value"] - b["config"]["rays"] / (b["ms_per_step"] * 1e-3) / 1e6) < 1e-6 * b["value"]
    e = b["e2e"]
    assert e["unit"] == b["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < b["value"] * 1.02  # the end-to-end figure is not a copy of the device-timed one
    assert b["gpu_launches"] > 0
    r = b["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] <= 1.0
    assert r["binding_limit"] and 0 < r["binding_frac"] <= 1.0
    c = b["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("port", "reference")
    cl = b["clocks"]
    assert not set(cl["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_line_matches_the_ours_line():
    b, r = load("r02_bench_n1.json"), load("r02_bench_reference_arm.json")
    assert r["impl"] == "reference"
    assert r["metric"] == b["metric"] and r["unit"] == b["unit"] and r["higher_is_better"] == b["higher_is_better"]
    assert r["config"]["workload"] == b["config"]["workload"]          # the driver compares these strings
    assert r["config"]["rays"] == b["config"]["rays"]
    assert r["config"]["mean_cells_per_ray"] == b["config"]["mean_cells_per_ray"]
    assert r["e2e"]["loss"] == b["e2e"]["loss"] == b["e2e_eager"]["loss"]   # same forward result in every arm and mode
    assert r["cpu_baseline"]["kind"] == "reference"
