"""bench.py without a GPU: the reference arm's JSON line (the CPU restatement timed on a bounded sample), the front-end
selection rule, and that the product arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def _run(*args):
    return subprocess.run([sys.executable, BENCH] + list(args), capture_output=True, text=True, timeout=600)


def test_reference_arm_line():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-streams", "2")
    assert p.returncode == 0, p.stderr[-500:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "PSS+SSS search Msamples/s" and line["unit"] == "Msamples/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 1
    assert line["config"]["workload"].startswith("C5 shard: 512 streams/GPU x 30.72 Msps fc32, D=16")
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_front_end_selection_rule():
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    assert bench.ntaps(16) == 525 and bench.ntaps(8) == 263 and bench.ntaps(1) == 0
    assert bench.tc_exists("fc32", 16) and bench.tc_exists("sc16", 4) and bench.tc_exists("sc8", 32)
    assert not bench.tc_exists("sc16", 2) and not bench.tc_exists("sc8", 4) and not bench.tc_exists("fc32", 20)
    argv = sys.argv
    try:
        picks = {}
        for key, extra in (("default", []), ("d4_fc32", ["--decim", "4"]), ("d4_sc16", ["--decim", "4", "--format", "sc16"]),
                           ("d5", ["--decim", "5"]), ("forced", ["--frontend", "fp32"]), ("c2", ["--workload", "c2"])):
            sys.argv = ["bench.py"] + extra
            picks[key] = bench.parse().frontend
    finally:
        sys.argv = argv
    # the tensor-core front end wherever it exists and wins; FP32 for fc32 at D <= 4, other rates, single-stream configs
    assert picks == {"default": "tc", "d4_fc32": "fp32", "d4_sc16": "tc", "d5": "fp32", "forced": "fp32", "c2": "fp32"}


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present: the refusal path is for boxes without one")
    p = _run("--steps", "1", "--warmup", "1", "--no-e2e")
    assert p.returncode != 0 and "no CUDA device" in (p.stderr + p.stdout)


def test_recorded_product_line_keeps_the_contract():
    """The product arm cannot run here (no device), so the line it printed on the B200 in the round's last run
    (profiles/bench_r02_v11_default.json) is checked instead: every key of the bench contract, the arithmetic that ties
    them together (value = samples per step / time, frac = achieved / peak, achieved = algorithmic bytes / kernel time),
    the bounds the cross-checks impose (kernel time below the step time, traffic within 1 % of the algorithmic bytes,
    end to end below the host link), parity and tolerance verdicts of the run."""
    with open(os.path.join(ROOT, "profiles", "bench_r02_v11_default.json")) as f:
        j = json.loads(f.read().strip().splitlines()[-1])
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert j["metric"] in base["metric"] and j["unit"] == "Msamples/s" and j["higher_is_better"] is True
    assert j["n_gpus"] == 1 and j["steps"] >= 10 and j["warmup"] >= 3 and j["scaling"] == "weak" and j["vs_baseline"] is None
    assert j["data"] == "synthetic" and "impl" not in j
    cfg = j["config"]
    assert cfg["workload"].startswith("C5 shard: 512 streams/GPU x 30.72 Msps fc32, D=16") and "model" not in cfg
    assert "larger than L2" in cfg["l2"]
    samples = cfg["streams_per_gpu"] * 30.72e6 * cfg["segment_ms"] / 1e3
    assert abs(j["value"] - samples / j["ms_per_step"] / 1e3) < 1e-6 * j["value"]
    ck = j["clocks"]
    assert ck["sm_mhz"] > 0.8 * ck["sm_max_mhz"] and not set(ck["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert j["gpu_launches"] >= j["steps"] * 4
    r = j["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["algorithmic_bytes_per_launch"] - samples * 8.5) < 1e-6 * samples          # 8 B read + 0.5 B written per sample
    t_kernel = r["stage_ms"]["frontend(convert+decimate)"]
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / t_kernel / 1e6) < 1e-6 * r["achieved"]
    assert t_kernel < j["ms_per_step"] and 0.99 < r["traffic"] / r["algorithmic_bytes_per_launch"] < 1.01
    assert 0.6 <= r["frac"] < 1.0 and r["kernel_alone"]["frac"] > r["frac"]
    e = j["e2e"]
    assert e["unit"] == "Msamples/s" and e["h2d_bytes_per_step"] == e["streams"] * 30.72e6 * cfg["segment_ms"] / 1e3 * 8
    assert e["d2h_bytes_per_step"] > 0 and e["value"] < j["value"] and e["host_link"]["frac_of_h2d_copy"] <= 1.0
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and 0 < cb["value"] < e["value"] and cb["sample"]
    assert j["parity_spot_check"]["bit_identical_to_oracle"] is True and j["parity_spot_check"]["ranks"] == 1
    assert j["tc_vs_fp32"]["decisions_identical"] is True and j["tc_vs_fp32"]["max_rel_diff_psr_peak"] < j["tc_vs_fp32"]["tolerance"] == 1e-4
    assert j["sustained"]["seconds"] >= 2.0 and j["sustained"]["value"] <= j["value"] * 1.02
