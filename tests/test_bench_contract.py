"""bench.py's contract, checked without a GPU: the host-side helpers (defaults, algorithmic-bytes formula of SURVEY §8(d),
ef calibration, the `config` object both arms print) and the JSON lines committed under profiles/ from the round's final
GPU runs — every key the driver reads is there and the numbers are consistent with each other."""
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)  # torch is imported inside main(), not here
    return mod


def _args(mod, *argv):
    old = sys.argv
    sys.argv = ["bench.py", *argv]
    try:
        return mod.parse_args()
    finally:
        sys.argv = old


def test_defaults_and_timing_rules():
    b = _bench()
    a = _args(b)
    assert (a.gpus, a.steps, a.warmup, a.impl) == (1, 20, 3, "islands_b200")
    assert (a.n, a.d, a.nq, a.dataset) == (1_000_000, 768, 10_000, "latent32")  # BASELINE configs[1]
    assert _args(b, "--warmup", "0", "--steps", "0").warmup == 3  # never fewer than three untimed steps
    assert _args(b, "--warmup", "0", "--steps", "0").steps == 1
    assert _args(b, "--island-nodes", "5000000", "--dim", "1024").n == 5_000_000  # the spellings torchrun leaves alone
    assert _args(b, "--impl", "reference").impl == "reference"


def test_algorithmic_bytes_formula():
    """B = n_dist·4d + n_edge·4 + n_hop·16 + nq·(4d + 12k)  (DESIGN §3)."""
    b = _bench()
    stats = np.array([[3, 100, 80, 0, 0], [5, 160, 120, 0, 0]], np.uint64)  # n_hop, n_edge, n_dist, n_adc, n_rerank
    total, per_query = b.algorithmic_bytes(stats, 768, 2, 10)
    assert total == 200 * 4 * 768 + 260 * 4 + 8 * 16 + 2 * (4 * 768 + 120)
    assert per_query == dict(n_hop=4.0, n_edge=130.0, n_dist=100.0)


def test_ef_calibration_picks_the_smallest_passing_ef():
    b = _bench()
    calls = []

    def recall_for(ef):
        calls.append(ef)
        return min(1.0, ef / 110.0)  # 0.95 is reached at ef = 104.5 -> first rung of 8 at or above: 112 on the coarse ladder 128

    ef, curve = b.calibrate_ef(recall_for, 0.95)
    assert ef == 112 and curve[ef] >= 0.95 and curve[104] < 0.95 and list(curve) == sorted(curve)
    assert calls[: calls.index(128) + 1] == [16, 24, 32, 48, 64, 96, 128]  # coarse ladder first, then the last interval
    ef, curve = b.calibrate_ef(recall_for, 0.95, fixed=64)
    assert ef == 64 and list(curve) == [64]
    ef, _ = b.calibrate_ef(lambda e: 0.1, 0.95)
    assert ef == b.EF_LADDER[-1]  # never reached: the largest rung is reported with its (insufficient) recall


def test_both_arms_print_the_same_config():
    b = _bench()
    a, r = _args(b), _args(b, "--impl", "reference")
    assert b.workload_config(a, 1_000_000, 768, 10_000, 1, 104) == b.workload_config(r, 1_000_000, 768, 10_000, 1, 104)
    cfg = b.workload_config(a, 1_000_000, 768, 10_000, 2, 104)
    assert cfg["total_nodes"] == 2_000_000 and "1000000 x 768" in cfg["workload"] and "larger than L2" in cfg["l2"]


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.load(f)


def test_committed_bench_line_carries_the_contract():
    line = _line("r02_bench_line.json")
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in line, key
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert line["metric"].split(" (")[0] in base["metric"] or base["metric"].split(" (")[0] in line["metric"]
    assert line["n_gpus"] == 1 and line["warmup"] >= 3 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["dtype"] == "f32" and line["data"] == "synthetic" and "workload" in line["config"] and "model" not in line["config"]
    # value = queries of the timed steps / device time; e2e is measured separately and is not the same number
    nq = 10_000
    assert abs(line["value"] - nq * 1000.0 / line["ms_per_step"]) / line["value"] < 1e-6
    e2e = line["e2e"]
    assert e2e["unit"] == line["unit"] and e2e["value"] != line["value"] and e2e["value"] < line["value"]
    assert e2e["h2d_bytes_per_step"] == nq * 768 * 4 and e2e["d2h_bytes_per_step"] == nq * (10 * 12 + 4)
    assert line["gpu_launches"] == line["steps"]  # one search kernel per step
    roof = line["roofline"]
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
    assert abs(roof["achieved"] - roof["algorithmic_bytes_per_launch"] / roof["kernel_ms"] / 1e6) / roof["achieved"] < 1e-6
    assert 0.5 < roof["frac"] < 1.0 and roof["kernel_ms"] <= line["ms_per_step"]
    assert 0.98 < roof["traffic"] / roof["algorithmic_bytes_per_launch"] < 1.10  # ncu DRAM bytes vs algorithmic bytes
    cpu = line["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["cores"] >= 1 and cpu["unit"] == line["unit"] and "ids equal to GPU: True" in cpu["sample"]
    clocks = line["clocks"]
    assert clocks["sm_mhz"] <= clocks["sm_max_mhz"] and not set(clocks["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_committed_reference_line_carries_the_contract():
    ref, ours = _line("r02_bench_reference_line.json"), _line("r02_bench_line.json")
    assert ref["impl"] == "reference" and ref["gpu_launches"] == 0
    for key in ("metric", "unit", "higher_is_better", "config"):
        assert ref[key] == ours[key], key  # same metric, same workload object: the driver's same_config holds
    assert ref["cpu_baseline"]["value"] == ref["value"] and ref["cpu_baseline"]["kind"] == "port"
    assert ref["e2e"] == {"value": ref["value"], "unit": ref["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert 50 < ours["e2e"]["value"] / ref["value"] < 200  # the headline ratio of the round (99 x on this box)
