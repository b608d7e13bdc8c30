"""CPU suite: the committed bench lines (profiles/) carry every key of the bench contract, so that a change to
bench.py that drops one is caught without a GPU."""
import json
import os

import pytest

from conftest import ROOT

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"]


def _load(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not committed")
    with open(path) as f:
        return json.loads([l for l in f.read().splitlines() if l.startswith("{")][-1])


def test_ours_line_has_the_contract_keys():
    d = _load("r01_bench_n1.json")
    for k in REQUIRED + ["cpu_baseline"]:
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["unit"] == "Mpixels/s" and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] == 24 * 4096 * 4096 and d["e2e"]["value"] < d["value"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel"):
        assert k in d["roofline"], k
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    assert d["cpu_baseline"]["kind"] in ("reference", "port")
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"], k
    assert d["gpu_launches"] > 0


def test_reference_line_has_the_contract_keys():
    d = _load("r01_bench_reference_n1.json")
    assert d["impl"] == "reference" and d["unit"] == "Mpixels/s" and d["higher_is_better"] is True
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for k in ("kind", "cores", "sample", "value"):
        assert k in d["cpu_baseline"], k


def test_multi_gpu_lines():
    d = _load("r01_bench_n2.json")
    assert d["n_gpus"] == 2 and d["scaling"] == "weak" and "replicas" in d["config"]["parallelism"]
    s = _load("r01_bench_shard_n2.json")
    assert s["n_gpus"] == 2 and s["scaling"] == "strong" and "chain-sharded" in s["config"]["parallelism"]
