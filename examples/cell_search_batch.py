#!/usr/bin/env python
"""cell_search_batch.py -- search many IQ captures for LTE cells in one batched GPU engine.

The multi-file form of examples/cell_search_file.py (reference: examples/cell_search_file.py:33-204,
one flowgraph per file): every file is one stream of a single `Trigger` engine, all streams advance
together in 50 ms passes, and the host decodes the PBCH of tagged half-frames until each stream has
published its cell -- what one `downlink_trigger_c -> cellstore` per file does, for N files at the
cost of one.  Output: one JSON object per file, the reference's cell dictionary plus "file" and
"status" ("FOUND" / "NOT_FOUND").

  python examples/cell_search_batch.py -s 7.68M --repeat --time-out 1 capture_a.fc32 capture_b.fc32
  python examples/cell_search_batch.py -s 30.72M --format sc16 --cut-off 30.72M *.sc16
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

from cell_search_file import REQUIRED_SAMPLE_RATE, eng_float, eng_int   # noqa: E402

PHICH_RESOURCES = ("1/6", "1/2", "1", "2")                   # lib/mib_impl.cc pack_cell


def search(args):
    import ltetrigger_b200 as lt
    from ltetrigger_b200 import _abi as A
    if args.sample_rate % REQUIRED_SAMPLE_RATE:
        sys.stderr.write("Sample rate {:.2f} MHz is not a multiple of 1.92 MHz. "
                         "Arbitrary resampling not supported at this time.\n".format(args.sample_rate / 1e6))
        sys.exit(-1)
    decim = int(args.sample_rate / REQUIRED_SAMPLE_RATE)
    fmt = {"fc32": lt.FMT_FC32, "sc16": lt.FMT_SC16, "sc8": lt.FMT_SC8}[args.format]
    dtype = A.FMT_DTYPE[fmt]
    files = [np.fromfile(f, dtype) for f in args.filenames]
    if fmt != lt.FMT_FC32:
        files = [d[:len(d) // 2 * 2].reshape(-1, 2) for d in files]
    n_streams = len(files)
    chunk = 96000 * decim                                     # 50 ms per pass, like the single-file tool
    trig = lt.Trigger(n_streams=n_streams, decim=decim, psr_threshold=max(args.threshold, lt.MIN_PSR_THRESHOLD),
                      max_chunk=chunk, input_format=fmt, record_all=False, keep_halfframes=True,
                      corr_mode=lt.CORR_FFT)
    found = [None] * n_streams
    buf = np.zeros((n_streams, chunk) + ((2,) if fmt != lt.FMT_FC32 else ()), dtype)
    fed, t_start = 0, time.time()
    while not all(found):
        if args.cut_off > -1 and fed >= args.cut_off:
            break
        if args.cut_off == -1 and args.time_out > -1 and time.time() - t_start >= args.time_out:
            break
        live = False
        buf[...] = 0                                          # a file that has ended contributes silence
        for s, d in enumerate(files):
            pos = fed % len(d) if args.repeat else fed
            if pos >= len(d) or found[s]:
                continue
            take = min(chunk, len(d) - pos)
            buf[s, :take] = d[pos:pos + take]
            live = True
        if not live:
            break
        recs = trig.process(buf)
        emitted = recs[(recs["flags"] & lt.F_EMIT) != 0]
        if len(emitted):
            hfs = trig.fetch_halfframes(len(emitted))
            for rec, hf in zip(emitted, hfs):
                s = int(rec["stream"])
                if found[s] or not (rec["flags"] & lt.F_CELL):
                    continue
                m = A.Mib()
                cp_norm = int(bool(rec["flags"] & lt.F_CP_NORM))
                if A.lib().ltb_mib_decode(np.ascontiguousarray(hf).ctypes.data, int(rec["cell_id"]), cp_norm, C.byref(m)) == 1:
                    found[s] = {"cell_id": int(rec["cell_id"]), "nof_tx_ports": int(m.nof_ports),
                                "cp_len": "Normal" if cp_norm else "Extended", "nof_prb": int(m.nof_prb),
                                "phich_len": "Normal" if m.phich_length == 0 else "Extended",
                                "nof_phich_resources": PHICH_RESOURCES[m.phich_resources],
                                "sfn_offset": int(m.sfn) & ~3, "tracking_start_time": int(time.time())}
        fed += chunk
    trig.close()
    return found


def main(args):
    print("Starting cell search on %d files... " % len(args.filenames), end="")
    sys.stdout.flush()
    found = search(args)
    print("done.")
    results = []
    for fname, cell in zip(args.filenames, found):
        out = dict(cell) if cell else {}
        out["status"] = "FOUND" if cell else "NOT_FOUND"
        out["file"] = fname
        results.append(json.dumps(out, indent=4))
    for r in results:
        print(r)
    return results


def parse(argv=None):
    def filetype(fname):
        if os.path.isfile(fname):
            return fname
        raise argparse.ArgumentTypeError("file {} does not exist".format(fname))

    parser = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    parser.add_argument("filenames", type=filetype, nargs="+")
    parser.add_argument("-s", "--sample-rate", type=eng_float, required=True, metavar="Hz",
                        help="sample rate of every input file [Required]")
    parser.add_argument("--format", default="fc32", choices=["fc32", "sc16", "sc8"],
                        help="sample format of the files [default=%(default)s, the reference's]")
    parser.add_argument("--repeat", action="store_true", help="loop files until all cells found or cut-off reached")
    parser.add_argument("-c", "--cut-off", type=eng_int, metavar="N", default=-1, help="stop after N input-rate samples per file")
    parser.add_argument("--time-out", type=eng_float, metavar="sec", default=-1, help="max time in seconds to perform search")
    parser.add_argument("--threshold", type=eng_float, default=4, help="peak to side-lobe ratio threshold")
    return parser.parse_args(argv)


if __name__ == "__main__":
    main(parse())
