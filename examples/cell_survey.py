#!/usr/bin/env python
"""cell_survey.py -- list every LTE cell in one long IQ capture, the capture sharded by time segment.

examples/cell_search_file.py (reference: examples/cell_search_file.py:33-204) walks a capture front to back and stops at
the first cell.  This tool answers "which cells are in this recording, and from when": the capture is cut into N
overlapping time segments (shard.plan_time_segments: 120 ms of halo, the stretch a chain needs to reach tracking), the
segments run as the N streams of one GPU engine, the host decodes the PBCH of tagged half-frames until each cell is
confirmed, and the records are stitched back onto the capture's time axis.  Output: one JSON object per cell, the
reference's cell dictionary (lib/mib_impl.cc:185-251) plus "first_seen_s", "last_seen_s" and "halfframes".

  python examples/cell_survey.py -s 30.72M --format sc16 --segments 32 capture.sc16
  python examples/cell_survey.py --frontend tc recording.sigmf-meta          # rate and sample type from the metadata
"""
from __future__ import print_function

import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gr-ltetrigger_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "examples"))

from cell_search_file import REQUIRED_SAMPLE_RATE, eng_float   # noqa: E402
from cell_search_batch import PHICH_RESOURCES                  # noqa: E402


def survey(args):
    import ltetrigger_b200 as lt
    from ltetrigger_b200 import _abi as A, shard, sigmf
    rec = None
    if sigmf.is_sigmf(args.filename):
        try:
            rec = sigmf.load(args.filename)
        except sigmf.SigMFError as e:
            sys.stderr.write("%s\n" % e)
            sys.exit(-1)
        args.sample_rate = rec["sample_rate"]
    if args.sample_rate is None:
        sys.stderr.write("-s / --sample-rate is required unless the file is a SigMF recording.\n")
        sys.exit(-1)
    if args.sample_rate % REQUIRED_SAMPLE_RATE:
        sys.stderr.write("Sample rate {:.2f} MHz is not a multiple of 1.92 MHz. "
                         "Arbitrary resampling not supported at this time.\n".format(args.sample_rate / 1e6))
        sys.exit(-1)
    decim = int(args.sample_rate / REQUIRED_SAMPLE_RATE)
    if rec is not None:
        fmt, iq = rec["input_format"], rec["samples"]
    else:
        fmt = {"fc32": lt.FMT_FC32, "sc16": lt.FMT_SC16, "sc8": lt.FMT_SC8}[args.format]
        iq = np.fromfile(args.filename, A.FMT_DTYPE[fmt])
        if fmt != lt.FMT_FC32:
            iq = iq[:len(iq) // 2 * 2].reshape(-1, 2)
    tc = args.frontend == "tc"
    if tc and fmt == lt.FMT_FC32 and not args.full_scale:
        sys.stderr.write("--frontend tc on fc32 input needs --full-scale (the range of the source, e.g. 1.0).\n")
        sys.exit(-1)
    plan = shard.plan_time_segments(len(iq), decim, args.segments, halo_halfframes=args.halo)
    rows = shard.cut_segments(iq, plan)
    chunk = 96000 * decim
    trig = lt.Trigger(n_streams=plan.n_segments, decim=decim, psr_threshold=max(args.threshold, lt.MIN_PSR_THRESHOLD),
                      max_chunk=chunk, input_format=fmt, keep_halfframes=True, corr_mode=lt.CORR_FFT,
                      frontend_mode=lt.FRONTEND_TC_INT if tc else lt.FRONTEND_FP32,
                      fc32_full_scale=args.full_scale if (tc and fmt == lt.FMT_FC32) else 0.0)
    t0 = time.time()
    recs, mibs = [], {}
    for a in range(0, plan.length, chunk):
        r = trig.process(rows[:, a:a + chunk])
        recs.append(r)
        emitted = r[(r["flags"] & lt.F_EMIT) != 0]
        if not len(emitted):
            continue
        for w, hf in zip(emitted, trig.fetch_halfframes(len(emitted))):
            cell = int(w["cell_id"])
            if not (w["flags"] & lt.F_CELL) or cell in mibs:
                continue
            m = A.Mib()
            cp_norm = int(bool(w["flags"] & lt.F_CP_NORM))
            if A.lib().ltb_mib_decode(np.ascontiguousarray(hf).ctypes.data, cell, cp_norm, C.byref(m)) == 1:
                mibs[cell] = {"cell_id": cell, "nof_tx_ports": int(m.nof_ports),
                              "cp_len": "Normal" if cp_norm else "Extended", "nof_prb": int(m.nof_prb),
                              "phich_len": "Normal" if m.phich_length == 0 else "Extended",
                              "nof_phich_resources": PHICH_RESOURCES[m.phich_resources],
                              "sfn_offset": int(m.sfn) & ~3}
    trig.close()
    st = shard.stitch_segments(np.concatenate(recs), plan)
    cells = []
    for d in shard.detections(st):
        cell = int(d["cell_id"])
        if cell not in mibs:                                  # tagged by sss but never confirmed by a MIB: not a cell
            continue
        g = st[((st["flags"] & lt.F_CELL) != 0) & (st["cell_id"] == cell)]
        out = dict(mibs[cell])
        out.update(first_seen_s=float(g["emit_start"].min() / REQUIRED_SAMPLE_RATE),
                   last_seen_s=float(g["emit_start"].max() / REQUIRED_SAMPLE_RATE), halfframes=int(len(g)))
        cells.append(out)
    info = {"segments": plan.n_segments, "segment_s": plan.length / args.sample_rate,
            "capture_s": len(iq) / args.sample_rate, "search_wall_s": time.time() - t0}
    return cells, info


def main(args):
    cells, info = survey(args)
    sys.stderr.write("%d segments of %.3f s over %.3f s of signal, searched in %.3f s\n"
                     % (info["segments"], info["segment_s"], info["capture_s"], info["search_wall_s"]))
    if not cells:
        print("NOT_FOUND")
    for c in cells:
        print(json.dumps(c, indent=4))
    return cells


def parse(argv=None):
    def filetype(fname):
        if os.path.isfile(fname) or os.path.isfile(fname + ".sigmf-meta"):
            return fname
        raise argparse.ArgumentTypeError("file {} does not exist".format(fname))

    parser = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    parser.add_argument("filename", type=filetype)
    parser.add_argument("-s", "--sample-rate", type=eng_float, default=None, metavar="Hz",
                        help="sample rate of the capture (a multiple of 1.92 MHz); required unless the file is a SigMF recording")
    parser.add_argument("--format", default="fc32", choices=["fc32", "sc16", "sc8"],
                        help="sample format of the file [default=%(default)s, the reference's]")
    parser.add_argument("--segments", type=int, default=16, metavar="N",
                        help="time segments searched side by side [default=%(default)s; fewer if the capture is short]")
    parser.add_argument("--halo", type=int, default=24, metavar="HALF_FRAMES",
                        help="overlap of consecutive segments in 5 ms half-frames [default=%(default)s]")
    parser.add_argument("--threshold", type=eng_float, default=4, help="peak to side-lobe ratio threshold")
    parser.add_argument("--frontend", default="fp32", choices=["fp32", "tc"],
                        help="resampler arithmetic: canonical float32 (default) or exact integers on the tensor cores")
    parser.add_argument("--full-scale", type=float, default=0.0, metavar="A",
                        help="with --frontend tc on fc32 input: the range of the source (23-bit fixed point over +-A)")
    return parser.parse_args(argv)


if __name__ == "__main__":
    main(parse())
