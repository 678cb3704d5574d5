"""Writes tests/golden/fixture_traces.json: the oracle's per-window trace of the four
bundled test_frames (0.5 s each, threshold 4).  Run from the repo root:
    python tests/golden/make_golden.py
Provenance: produced by oracle/ltetrigger_oracle.c (restated srsLTE release_18_06_1
semantics; the reference itself cannot be run offline).  The only quantities in it that
the reference's own tests pin are cell_id and cp type."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from conftest import FIXTURES, load_fixture  # noqa: E402
from oracle import oracle as O  # noqa: E402

out = {}
for name in FIXTURES:
    x, decim, cell_id = load_fixture(name, 0.5)
    recs = O.trigger_run(x[None, :], decim=decim)
    out[name] = {"n_records": len(recs), "decim": decim, "cell_id": cell_id}
    for field in ("win_start", "emit_start", "flags", "peak_pos", "score", "m0", "m1", "cell_id"):
        out[name][field] = recs[field].tolist()
    out[name]["psr_bits"] = recs["psr"].view(np.uint32).tolist()
    out[name]["cfo_bits"] = recs["cfo"].view(np.uint32).tolist()
    # the overlap-save evaluation of the matched filter (ORC_CONV_OS): same decisions, own PSR bits
    os_recs = O.trigger_run(x[None, :], decim=decim, conv_mode=O.CONV_OS)
    assert all((os_recs[f] == recs[f]).all() for f in ("win_start", "emit_start", "flags", "peak_pos", "m0", "m1", "cell_id"))
    out[name]["os_psr_bits"] = os_recs["psr"].view(np.uint32).tolist()
    out[name]["os_peak_value_bits"] = os_recs["peak_value"].view(np.uint32).tolist()
with open(os.path.join(HERE, "fixture_traces.json"), "w") as f:
    json.dump(out, f)
print("wrote", {k: v["n_records"] for k, v in out.items()})
