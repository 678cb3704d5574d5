"""Writes tests/golden/fixture_traces_tcint.json: the oracle's per-window trace of the bundled test_frames that have a
decimating front end (25 / 50 / 100 PRB at D = 4 / 8 / 16) with the integer front end (ORC_FRONT_TCINT), for the three
input formats: fc32 taken as fixed point over +-4.0 (the frames' peak is 2.6), sc16 and sc8 quantised by synth.to_sc16 /
to_sc8 (full scale 4.0).  Run from the repo root:
    python tests/golden/make_golden_tcint.py
Provenance: produced by oracle/ltetrigger_oracle.c; it freezes the integer arithmetic contract (tap quantisation, the
fc32 fixed-point grid, the one rounding per output) so that a change to it shows up on the CPU.  What the reference's
own tests pin in it: cell_id and cp type."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, "..", "..", "gr-ltetrigger_b200", "python"))
from conftest import load_fixture  # noqa: E402
from oracle import oracle as O  # noqa: E402
from ltetrigger_b200 import synth  # noqa: E402

FULL_SCALE = 4.0


def cases():
    for name in ("25prb", "50prb", "100prb"):
        x, decim, cell_id = load_fixture(name, 0.3)
        x = x[:len(x) // (8 * decim) * (8 * decim)]
        for fmt, iq in ((0, x[None, :]), (1, synth.to_sc16(x[None, :])), (2, synth.to_sc8(x[None, :]))):
            if fmt == 2 and decim < 8:
                continue                              # no sc8 kernel below D = 8
            yield "%s_fmt%d" % (name, fmt), iq, decim, fmt, cell_id


def trace(iq, decim, fmt):
    return O.trigger_run(iq, decim=decim, fmt=fmt, conv_mode=O.CONV_OS | O.FRONT_TCINT, fc32_full_scale=FULL_SCALE if fmt == 0 else 0.0)


if __name__ == "__main__":
    out = {}
    for key, iq, decim, fmt, cell_id in cases():
        recs = trace(iq, decim, fmt)
        out[key] = {"n_records": len(recs), "decim": decim, "fmt": fmt, "cell_id": cell_id,
                    "psr_bits": recs["psr"].view(np.uint32).tolist(), "peak_value_bits": recs["peak_value"].view(np.uint32).tolist(),
                    "cfo_bits": recs["cfo"].view(np.uint32).tolist()}
        for field in ("win_start", "emit_start", "flags", "peak_pos", "m0", "m1", "cell_id"):
            out[key][field] = recs[field].tolist()
    with open(os.path.join(HERE, "fixture_traces_tcint.json"), "w") as f:
        json.dump(out, f)
    print("wrote", {k: v["n_records"] for k, v in out.items()})
