#!/usr/bin/env python
"""Randomised parity campaign (a script, not a pytest module): seeded random engine configurations --
rate, wire format, front end (canonical FP32 or, where it exists, the integer tensor-core one with a random
fixed-point range for fc32), correlator, frame type, stream count, chunking, threshold, tracking parameters,
SNR, CP type, carrier offset -- each run through the C ABI in ragged chunks and compared record for
record, bit for bit, with the CPU oracle.  Prints one line per case and exits non-zero on the first
mismatch with the configuration that produced it.

    python tests/fuzz_parity.py --seconds 300 --seed 1
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gr-ltetrigger_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--max-cases", type=int, default=100000)
    a = ap.parse_args()
    n_cases, bad = campaign(a.seconds, a.seed, a.max_cases)
    if bad is not None:
        sys.exit(1)
    print("fuzz_parity: %d cases bit-identical to the oracle" % n_cases)


def campaign(seconds, seed, max_cases=100000, log=print):
    """Run random cases for `seconds`; returns (cases run, None) or (cases run, (config, error)) at the
    first mismatch.  tests/test_gpu_campaigns.py runs a 60 s campaign inside `pytest -m gpu`."""
    import ltetrigger_b200 as lt
    from ltetrigger_b200 import synth
    from oracle import oracle as O
    from conftest import assert_recs_equal

    rng = np.random.default_rng(seed)
    t_end = time.time() + seconds
    n_cases = 0
    while time.time() < t_end and n_cases < max_cases:
        decim = int(rng.choice([1, 1, 2, 3, 4, 5, 6, 8, 8, 10, 12, 13, 15, 16, 16, 20, 24, 32]))
        fmt = int(rng.integers(0, 3))
        corr = int(rng.integers(0, 2))
        tdd = int(rng.integers(0, 4) == 0)
        n_streams = int(rng.integers(1, 5))
        thr = float(rng.choice([1.7, 2.5, 4.0, 4.0]))
        track_after = int(rng.choice([4, 16, 16]))
        track_every = int(rng.choice([2, 8, 8]))
        n_search = int(rng.integers(30, 70)) * 4800                 # 75 .. 175 ms of search-rate samples
        n = n_search * decim
        step = 8 * decim
        chunk = int(rng.integers(20, 6000)) * step
        # the integer tensor-core front end where the kernel exists (half of those cases); fc32 gets a declared range
        # between 3 x and 40 x the signal's rms, so some cases clip and others use few of the grid's bits
        tc_ok = decim in {0: (2, 4, 8, 12, 16, 24, 32), 1: (4, 8, 12, 16, 24, 32), 2: (8, 16, 24, 32)}[fmt]
        tc = bool(tc_ok and rng.integers(0, 2))
        fs = float(rng.uniform(3.0, 40.0)) if (tc and fmt == 0) else 0.0
        cfg = dict(decim=decim, fmt=fmt, corr=corr, tdd=tdd, n_streams=n_streams, thr=thr, track_after=track_after,
                   track_every=track_every, n=n, chunk=chunk, tc=tc, fs=fs)
        rows = []
        for s in range(n_streams):
            kind = rng.integers(0, 6)
            cell = int(rng.integers(0, 504))
            if kind == 0:                                           # noise only / silence
                x = synth.capture(cell, n, snr_db=0.0, decim=decim, seed=int(rng.integers(1 << 30)), noise_only=True)
                if rng.integers(0, 2):
                    x[:] = 0
            else:
                x = synth.capture(cell, n, snr_db=float(rng.uniform(-6, 15)), decim=decim, seed=int(rng.integers(1 << 30)),
                                  cfo_hz=float(rng.choice([0.0, 0.0, rng.uniform(-6000, 6000)])),
                                  ext_cp=bool(rng.integers(0, 4) == 0), tdd=bool(tdd))
            rows.append(x)
        x = np.stack(rows)
        iq = x if fmt == 0 else (synth.to_sc16(x) if fmt == 1 else synth.to_sc8(x))
        conv = (O.CONV_OS if corr else O.CONV_DIRECT) | (O.FRAME_TDD if tdd else 0) | (O.FRONT_TCINT if tc else 0)
        want = O.trigger_run(iq, decim=decim, fmt=fmt, psr_threshold=thr, track_after=track_after, track_every=track_every,
                             conv_mode=conv, fc32_full_scale=fs)
        trig = lt.Trigger(n_streams=n_streams, decim=decim, psr_threshold=thr, max_chunk=chunk, input_format=fmt,
                          track_after=track_after, track_every=track_every, corr_mode=corr, frame_type=tdd,
                          frontend_mode=lt.FRONTEND_TC_INT if tc else lt.FRONTEND_FP32, fc32_full_scale=fs)
        got = trig.run(iq, chunk=chunk)
        trig.close()
        try:
            assert_recs_equal(got, want)
        except AssertionError as e:
            log("MISMATCH %r %s" % (cfg, e))
            return n_cases, (cfg, str(e))
        n_cases += 1
        cells = int(((got["flags"] & lt.F_CELL) != 0).sum())
        log("ok %4d D=%-2d fmt=%d fe=%s corr=%d tdd=%d S=%d thr=%.1f ta=%d te=%d chunk=%d recs=%d tagged=%d" % (
            n_cases, decim, fmt, ("tc(%.1f)" % fs if fmt == 0 else "tc") if tc else "fp32", corr, tdd, n_streams, thr, track_after,
            track_every, chunk, len(got), cells))
    return n_cases, None


if __name__ == "__main__":
    main()
