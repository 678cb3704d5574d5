"""One long capture, sequential against time-segment sharded (shard.plan_time_segments), on one GPU.

The reference walks a capture front to back in one flowgraph (examples/cell_search_file.py:56-60); with one stream the
GPU engine runs three chains on 148 SMs (BASELINE configs c1-c3, tools/bench_single.py).  Cut into N overlapping time
segments the same capture fills the machine: the segments are the N streams of one engine, read in place from ONE pinned
host buffer (row k starts k * step samples in, the rows overlap by the halo; every pass copies a 50 ms column of all rows
with one pitched copy), and the records are stitched back onto the capture's time axis.  Timed per format: wall time
from the first submit to the last collect, host buffers, H2D inside -- sequential (one stream, 100 ms calls, two in
flight) and N = 8 / 32 / 128 segments; the stitched list is compared with the sequential one (tagged half-frames)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gr-ltetrigger_b200", "python"))


def run(seconds=8.0, fmt_name="sc16", segments=(8, 32, 128), device=0, threshold=4.0):
    import torch
    import ltetrigger_b200 as lt
    from ltetrigger_b200 import shard, synth, _abi as A
    decim = 16
    fmt = {"fc32": lt.FMT_FC32, "sc16": lt.FMT_SC16}[fmt_name]
    bps = A.FMT_BYTES[fmt]
    frame = np.fromfile(os.path.join(ROOT, "tests", "golden", "test_frames", "lte_frame_100prb_cellid_369"), np.complex64)
    gran = 8 * decim * 128                                   # every plan below has uniform starts on this capture
    n = int(round(seconds * 30.72e6)) // gran * gran
    x = np.tile(frame, -(-n // len(frame)))[:n]
    host = torch.empty((n * bps,), dtype=torch.uint8, pin_memory=True)
    src = x if fmt == lt.FMT_FC32 else synth.to_sc16(x[None, :])[0]
    host.numpy()[:] = np.ascontiguousarray(src).view(np.uint8).reshape(-1)
    base = host.data_ptr()
    out = {"capture_s": n / 30.72e6, "format": fmt_name, "decim": decim, "fixture": "lte_frame_100prb_cellid_369 tiled",
           "timed": "first submit to last collect, pinned host buffer, H2D inside; median of 3"}

    def timed(n_streams, stride, length, chunk):
        eng = lt.Trigger(n_streams=n_streams, decim=decim, psr_threshold=threshold, max_chunk=chunk, input_format=fmt,
                         device=device, corr_mode=lt.CORR_FFT, frontend_mode=lt.FRONTEND_TC_INT if fmt != lt.FMT_FC32 else lt.FRONTEND_FP32)
        offs = list(range(0, length, chunk))
        best, recs = [], None
        for rep in range(4):
            eng.reset()
            got = []
            t0 = time.perf_counter()
            eng.submit_host_ptr(base + offs[0] * bps, stride * bps, min(chunk, length - offs[0]))
            for a in offs[1:]:
                eng.submit_host_ptr(base + a * bps, stride * bps, min(chunk, length - a))
                got.append(eng.collect().copy())
            got.append(eng.collect().copy())
            if rep:
                best.append(time.perf_counter() - t0)
            recs = np.concatenate(got)
        eng.close()
        return sorted(best)[1], recs

    dt, seq = timed(1, n, n, 192000 * decim)
    out["sequential"] = {"wall_ms": 1e3 * dt, "msamples_per_s": n / dt / 1e6, "realtime_factor": n / dt / 30.72e6}
    cells_seq = np.sort(seq["emit_start"][(seq["flags"] & lt.F_CELL) != 0])
    out["segmented"] = []
    for want in segments:
        plan = shard.plan_time_segments(n, decim, want)
        step = int(plan.starts[1] - plan.starts[0]) if plan.n_segments > 1 else n
        uniform = plan.n_segments > 1 and bool((np.diff(plan.starts) == step).all())
        if not uniform:
            out["segmented"].append({"segments": plan.n_segments, "skipped": "non-uniform starts"})
            continue
        chunk = min(96000 * decim, step) // (8 * decim) * (8 * decim)
        dt, recs = timed(plan.n_segments, step, plan.length, chunk)
        st = shard.stitch_segments(recs, plan)
        cells = np.sort(st["emit_start"][(st["flags"] & lt.F_CELL) != 0])
        out["segmented"].append({"segments": plan.n_segments, "segment_s": plan.length / 30.72e6, "halo_s": plan.halo / 30.72e6,
                                 "wall_ms": 1e3 * dt, "msamples_per_s": n / dt / 1e6, "realtime_factor": n / dt / 30.72e6,
                                 "samples_read_over_capture": plan.n_segments * plan.length / n,
                                 "speedup_vs_sequential": out["sequential"]["wall_ms"] / (1e3 * dt),
                                 "tagged_halfframes": int(len(cells)), "tagged_halfframes_sequential": int(len(cells_seq)),
                                 "same_tagged_halfframes": bool(np.array_equal(cells, cells_seq)),
                                 "cells": sorted(set(st["cell_id"][(st["flags"] & lt.F_CELL) != 0].tolist()))})
    return out


if __name__ == "__main__":
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
    for f in (sys.argv[2:] or ["sc16", "fc32"]):
        print(json.dumps(run(secs, f)))
