#!/usr/bin/env python
"""Detection rate against SNR on batched synthetic captures (BASELINE config C4).

The batch counterpart of the reference's examples/snr_ltetrigger.grc (a test frame times a gain
plus a Gaussian noise source into downlink_trigger_c, threshold 1.5..6 on a slider): here 256
seeded streams per SNR point with cell ids dealt from a permutation of 0..503 run through the
batched engine at once.  Per SNR point it reports how many streams ended with the right
cell_id (majority over the tagged half-frames) and how many tagged a wrong one.
tests/c4_snr_sweep.py runs the same sweep with every record checked against the CPU oracle.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gr-ltetrigger_b200", "python"))


def _one(args):
    from ltetrigger_b200 import synth
    cell, n, snr, seed = args
    return synth.capture(cell, n, snr, 1, seed=seed)


def make_batch(pool, n_streams, n, snr, master_seed):
    perm = np.random.default_rng(master_seed).permutation(504)
    ids = np.array([perm[i % 504] for i in range(n_streams)])
    jobs = [(int(ids[i]), n, snr, master_seed ^ i) for i in range(n_streams)]
    rows = pool.map(_one, jobs, chunksize=4) if pool else [_one(j) for j in jobs]
    return np.stack(rows), ids


def main(check=None):
    """check(iq, records, threshold, corr) -> dict merged into the SNR point (used by tests/c4_snr_sweep.py)."""
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--streams", type=int, default=256)
    ap.add_argument("--seconds", type=float, default=0.5)
    ap.add_argument("--snr-min", type=int, default=-10)
    ap.add_argument("--snr-max", type=int, default=10)
    ap.add_argument("--snr-step", type=int, default=1)
    ap.add_argument("-t", "--threshold", type=float, default=4.0)
    ap.add_argument("--seed", type=int, default=20260)
    ap.add_argument("--corr", default="fft", choices=["fft", "direct"], help="matched-filter evaluation")
    ap.add_argument("--workers", type=int, default=os.cpu_count())
    ap.add_argument("-o", "--output", default=None)
    a = ap.parse_args()

    import ltetrigger_b200 as lt
    n = int(a.seconds * 1.92e6) // 8 * 8
    pool = mp.get_context("fork").Pool(a.workers) if a.workers > 1 else None
    trig = lt.Trigger(n_streams=a.streams, decim=1, psr_threshold=a.threshold, max_chunk=n,
                      corr_mode=lt.CORR_FFT if a.corr == "fft" else lt.CORR_DIRECT)
    points = []
    for snr in range(a.snr_min, a.snr_max + 1, a.snr_step):
        iq, ids = make_batch(pool, a.streams, n, float(snr), a.seed + 1000 * (snr + 100))
        trig.reset()
        t0 = time.perf_counter()
        got = trig.run(iq)
        dt = time.perf_counter() - t0
        right = wrong = 0
        first = []
        tagged = got[(got["flags"] & lt.F_CELL) != 0]
        for s in range(a.streams):
            c = tagged[tagged["stream"] == s]
            if len(c) == 0:
                continue
            if np.bincount(c["cell_id"]).argmax() == ids[s]:
                right += 1
                first.append(int(c["emit_start"].min()))
            else:
                wrong += 1
        pt = {"snr_db": snr, "streams": a.streams, "detected": right, "wrong_cell": wrong,
              "detection_rate": right / a.streams, "records": int(len(got)),
              "median_first_tag_ms": (float(np.median(first)) / 1920.0 if first else None),
              "engine_wall_ms": 1e3 * dt}
        if check is not None:
            pt.update(check(iq, got, a.threshold, a.corr))
        points.append(pt)
        print(json.dumps(pt), flush=True)
    if pool:
        pool.close()
    out = {"config": "C4: %d streams x %.2f s at 1.92 Msps per SNR point, threshold %.1f, seed %d, %s correlator" % (
        a.streams, a.seconds, a.threshold, a.seed, a.corr), "points": points}
    if a.output:
        with open(a.output, "w") as f:
            json.dump(out, f, indent=1)
    return points


if __name__ == "__main__":
    main()
