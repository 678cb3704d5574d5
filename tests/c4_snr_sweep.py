#!/usr/bin/env python
"""BASELINE config C4 at full size with the oracle as checker: examples/snr_sweep.py (the product's
batched sweep) plus, per SNR point, a bit-for-bit comparison of every window record with the CPU
oracle's.  A script, not a pytest module (it takes minutes); the reduced version that runs in
the GPU suite is test_gpu_parity.py::test_synthetic_snr_sweep_batched.

    python tests/c4_snr_sweep.py -o gpurun_out/snr_sweep.json
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))


def check(iq, got, threshold, corr):
    from oracle import oracle as O
    want = O.trigger_run(iq, decim=1, psr_threshold=threshold, conv_mode=O.CONV_OS if corr == "fft" else O.CONV_DIRECT)
    same = len(want) == len(got)
    for f in (want.dtype.names if same else ()):
        g, w = got[f], want[f]
        if g.dtype.kind == "f":                          # floats as bit patterns, +0 == -0
            same &= bool(((g.view(np.uint32) == w.view(np.uint32)) | ((g == 0) & (w == 0))).all())
        else:
            same &= bool((g == w).all())
    return {"records_bit_identical_to_oracle": bool(same)}


if __name__ == "__main__":
    import snr_sweep
    points = snr_sweep.main(check)
    sys.exit(0 if all(p["records_bit_identical_to_oracle"] for p in points) else 1)
