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
