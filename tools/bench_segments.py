"""One long capture, sequential against time-segment sharded (shard.plan_time_segments), on one GPU and, under
torchrun, across GPUs.

The reference walks a capture front to back in one flowgraph (examples/cell_search_file.py:56-60); with one stream the
GPU engine runs three chains on 148 SMs (BASELINE configs c1-c3, tools/bench_single.py).  Cut into N overlapping time
segments the same capture fills the machine: the segments are the N streams of one engine, read in place from ONE pinned
host buffer (row k starts k * step samples in, the rows overlap by the halo; every pass copies a 50 ms column of all rows
with one pitched copy), and the records are stitched back onto the capture's time axis.  Timed per format: wall time
from the first submit to the last collect, host buffers, H2D inside -- sequential (one stream, 100 ms calls, two in
flight) and N = 4 ... 32 segments; the stitched list is compared with the sequential one (tagged half-frames).

What it shows (profiles/bench_segments_r02.json): from host memory one stream already runs at the host link's rate
(8 s of 30.72 Msps sc16 in 21.7 ms, 45 GB/s), so on ONE GPU segments buy nothing end to end and the halo's re-reads
cost; the axis is for several GPUs, where every rank pulls its stretch of the capture over its own link (`run_dist`)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gr-ltetrigger_b200", "python"))


def run(seconds=8.0, fmt_name="sc16", segments=(4, 8, 16, 32), device=0, threshold=4.0):
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


def run_dist(seconds=8.0, fmt_name="sc16", segments_per_rank=1, threshold=4.0):
    """Under torchrun: the capture's segments dealt to the ranks in contiguous runs, every rank reading ITS stretch of
    the capture from its own pinned buffer over its own host link; barrier + max over ranks; records merged on rank 0
    (shard.merge_records over NCCL), stitched, and checked for gaps (tagged half-frames exactly 9600 samples apart)."""
    import torch
    import torch.distributed as dist
    import ltetrigger_b200 as lt
    from ltetrigger_b200 import shard, synth, _abi as A
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    decim = 16
    fmt = {"fc32": lt.FMT_FC32, "sc16": lt.FMT_SC16}[fmt_name]
    bps = A.FMT_BYTES[fmt]
    frame = np.fromfile(os.path.join(ROOT, "tests", "golden", "test_frames", "lte_frame_100prb_cellid_369"), np.complex64)
    gran = 8 * decim * 128
    n = int(round(seconds * 30.72e6)) // gran * gran
    plan = shard.plan_time_segments(n, decim, world * segments_per_rank)
    step = int(plan.starts[1] - plan.starts[0])
    assert plan.n_segments == world * segments_per_rank and bool((np.diff(plan.starts) == step).all())
    k = segments_per_rank
    s0 = int(plan.starts[rank * k])
    mine = (k - 1) * step + plan.length
    reps = -(-(s0 % len(frame) + mine) // len(frame))
    x = np.tile(frame, reps)[s0 % len(frame):s0 % len(frame) + mine]      # the fixture is periodic: this IS capture[s0 : s0 + mine]
    host = torch.empty((mine * bps,), dtype=torch.uint8, pin_memory=True)
    host.numpy()[:] = np.ascontiguousarray(x if fmt == lt.FMT_FC32 else synth.to_sc16(x[None, :])[0]).view(np.uint8).reshape(-1)
    base = host.data_ptr()
    chunk = (192000 * decim if k == 1 else min(96000 * decim, step)) // (8 * decim) * (8 * decim)
    eng = lt.Trigger(n_streams=k, decim=decim, psr_threshold=threshold, max_chunk=chunk, input_format=fmt, device=local,
                     corr_mode=lt.CORR_FFT, frontend_mode=lt.FRONTEND_TC_INT if fmt != lt.FMT_FC32 else lt.FRONTEND_FP32)
    offs = list(range(0, plan.length, chunk))
    times = []
    for rep in range(4):
        eng.reset()
        got = []
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        eng.submit_host_ptr(base + offs[0] * bps, max(step, plan.length) * bps if k == 1 else step * bps, min(chunk, plan.length - offs[0]))
        for a in offs[1:]:
            eng.submit_host_ptr(base + a * bps, max(step, plan.length) * bps if k == 1 else step * bps, min(chunk, plan.length - a))
            got.append(eng.collect().copy())
        got.append(eng.collect().copy())
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if rep:
            times.append(float(dt.item()))
    eng.close()
    recs = shard.to_global(np.concatenate(got), np.arange(rank * k, rank * k + k))
    merged = shard.merge_records(recs, dst=0)
    out = None
    if rank == 0:
        st = shard.stitch_segments(merged, plan)
        cells = np.sort(st["emit_start"][(st["flags"] & lt.F_CELL) != 0])
        dt = sorted(times)[1]
        out = {"n_gpus": world, "capture_s": n / 30.72e6, "format": fmt_name, "segments": plan.n_segments,
               "segment_s": plan.length / 30.72e6, "halo_s": plan.halo / 30.72e6, "wall_ms": 1e3 * dt,
               "msamples_per_s": n / dt / 1e6, "realtime_factor": n / dt / 30.72e6,
               "samples_read_over_capture": plan.n_segments * plan.length / n, "tagged_halfframes": int(len(cells)),
               "halfframe_spacing": sorted(set(np.diff(cells).tolist())),
               "cells": sorted(set(st["cell_id"][(st["flags"] & lt.F_CELL) != 0].tolist())),
               "timed": "barrier, first submit to last collect on every rank, max over ranks; pinned host buffers, H2D inside; median of 3"}
    dist.barrier()
    return out


if __name__ == "__main__":
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", os.environ["RANK"])))
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))   # once: re-initialising per configuration fails in NCCL's bootstrap
        for f in (sys.argv[2:] or ["sc16", "fc32"]):
            for spr in (1, 4):
                line = run_dist(float(sys.argv[1]) if len(sys.argv) > 1 else 8.0, f, spr)
                if line:
                    print(json.dumps(line), flush=True)
        dist.destroy_process_group()
        sys.exit(0)
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
    for f in (sys.argv[2:] or ["sc16", "fc32"]):
        print(json.dumps(run(secs, f)))
