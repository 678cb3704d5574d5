#!/usr/bin/env python
"""cell_search_file.py -- search an IQ capture for LTE cells on the GPU.

Same command line and output as the reference's examples/cell_search_file.py:33-204:
file_source(repeat) -> [rational_resampler_ccc(1, D)] -> [throttle] -> [head] ->
downlink_trigger_c(threshold, exit_on_success=True) -> cellstore, then one JSON object per found
cell ("status": "FOUND") or {"status": "NOT_FOUND"}, optionally also written to a FIFO as
"<length>\\n<json>".  The resampler is fused into the engine's front end; --throttle is accepted
and ignored (there is no CPU load to lower).

  python examples/cell_search_file.py tests/golden/test_frames/lte_frame_50prb_cellid_125 -s 15.36M --repeat --time-out 1

Beyond the reference: a SigMF recording (`name.sigmf-meta` + `name.sigmf-data`, datatype cf32_le, ci16_le or ci8) is
read with its own sample rate and type -- `-s` then is optional and must agree if given -- and integer types go to
the GPU as they are (the short-to-complex conversion is fused into the front end).
"""
from __future__ import print_function

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gr-ltetrigger_b200", "python"))

REQUIRED_SAMPLE_RATE = 1.92e6


def eng_float(s):
    """gnuradio.eng_arg.eng_float: 15.36M, 1.92e6, 30720k ..."""
    mult = {"k": 1e3, "M": 1e6, "G": 1e9, "m": 1e-3, "u": 1e-6}
    s = s.strip()
    if s and s[-1] in mult:
        return float(s[:-1]) * mult[s[-1]]
    return float(s)


def eng_int(s):
    return int(eng_float(s))


def search(args):
    import ltetrigger_b200 as lt
    from ltetrigger_b200 import sigmf
    input_format, data = lt.FMT_FC32, None
    if sigmf.is_sigmf(args.filename):
        try:
            rec = sigmf.load(args.filename)
        except sigmf.SigMFError as e:
            sys.stderr.write("%s\n" % e)
            sys.exit(-1)
        if args.sample_rate is not None and abs(args.sample_rate - rec["sample_rate"]) > 0.5:
            sys.stderr.write("Sample rate {:.6f} MHz given, but the SigMF metadata says {:.6f} MHz.\n".format(
                args.sample_rate / 1e6, rec["sample_rate"] / 1e6))
            sys.exit(-1)
        args.sample_rate, input_format, data = rec["sample_rate"], rec["input_format"], rec["samples"]
    elif args.sample_rate is None:
        sys.stderr.write("-s/--sample-rate is required for a raw fc32 file.\n")
        sys.exit(-1)
    if args.sample_rate % REQUIRED_SAMPLE_RATE:
        sys.stderr.write("Sample rate {:.2f} MHz is not a multiple of 1.92 MHz. "
                         "Arbitrary resampling not supported at this time.\n".format(args.sample_rate / 1e6))
        sys.exit(-1)
    decim = int(args.sample_rate / REQUIRED_SAMPLE_RATE)
    tc = args.frontend == "tc"
    if tc and input_format == lt.FMT_FC32 and not args.full_scale:
        sys.stderr.write("--frontend tc on fc32 input needs --full-scale (the range of the source, e.g. 1.0).\n")
        sys.exit(-1)
    trigger = lt.downlink_trigger_c(psr_threshold=args.threshold, exit_on_success=True, decim=decim, input_format=input_format,
                                    frontend_mode=lt.FRONTEND_TC_INT if tc else lt.FRONTEND_FP32,
                                    fc32_full_scale=args.full_scale if (tc and input_format == lt.FMT_FC32) else 0.0)
    store = lt.cellstore().connect(trigger)
    if data is None:
        data = np.fromfile(args.filename, np.complex64)
    chunk = 96000 * decim                                     # 50 ms per scheduler pass
    fed, t_start = 0, time.time()
    done = lambda: any(m.done for m in (trigger.mib0, trigger.mib1, trigger.mib2))   # WORK_DONE
    # the reference puts head(cut_off) AFTER the resampler (examples/cell_search_file.py:47-48, 69-77), so
    # -c counts search-rate samples: here the resampler is fused into the engine, i.e. cut_off * decim inputs
    cut_in = args.cut_off * decim if args.cut_off > -1 else -1
    while not done():
        if cut_in > -1 and fed >= cut_in:
            break
        if args.cut_off == -1 and args.time_out > -1 and time.time() - t_start >= args.time_out:
            break
        pos = fed % len(data) if args.repeat else fed
        if pos >= len(data):
            break                                             # end of file without --repeat
        take = min(chunk, len(data) - pos)
        if cut_in > -1:
            take = min(take, cut_in - fed)
        trigger.work(data[pos:pos + take])
        fed += take
    return store


def main(args):
    print("Starting cell search... ", end="")
    sys.stdout.flush()
    store = search(args)
    print("done.")
    results = []
    if store.tracking():
        for cell in store.cells():
            cell_json = dict(cell)
            cell_json["status"] = "FOUND"
            results.append(json.dumps(cell_json, indent=4))
    else:
        results.append(json.dumps({"status": "NOT_FOUND"}))
    for cell in results:
        print(cell)
    if args.fifoname:
        if not os.path.exists(args.fifoname):
            os.mkfifo(args.fifoname)
        pipeout = os.open(args.fifoname, os.O_WRONLY)
        for cell in results:
            os.write(pipeout, (str(len(cell)) + "\n" + cell).encode())
        os.close(pipeout)
    return results


def parse(argv=None):
    """The reference CLI's flags (examples/cell_search_file.py:140-200): same names, types and defaults."""
    def existing_file(name):
        if not os.path.isfile(name) and not os.path.isfile(name + ".sigmf-meta"):
            raise argparse.ArgumentTypeError("file {} does not exist".format(name))
        return name

    ap = argparse.ArgumentParser(description="search an IQ capture (fc32) for LTE cells on the GPU")
    ap.add_argument("filename", type=existing_file)
    options = [
        (("-s", "--sample-rate"), dict(type=eng_float, default=None, metavar="Hz", help="sample rate of the capture (a multiple of 1.92 MHz); required unless the file is a SigMF recording")),
        (("-f", "--frequency"), dict(type=eng_float, metavar="Hz", help="center frequency of the capture (informational)")),
        (("--repeat",), dict(action="store_true", help="start over at the end of the file until a cell is found, the cut-off or the time-out")),
        (("-c", "--cut-off"), dict(type=eng_int, metavar="N", default=-1, help="give up after N samples at the 1.92 Msps search rate (the reference counts them after the resampler)")),
        (("--throttle",), dict(type=eng_float, metavar="Hz", help="accepted for compatibility; the GPU path is not throttled")),
        (("--time-out",), dict(type=eng_float, metavar="sec", default=-1, help="give up after this many seconds")),
        (("--threshold",), dict(type=eng_float, default=4, help="PSR threshold of the trigger (clamped to > 1.5)")),
        (("--gui",), dict(action="store_true", help=argparse.SUPPRESS)),
        (("--debug",), dict(action="store_true", help=argparse.SUPPRESS)),
        (("--fifoname",), dict(default=None, help="also write every result to this FIFO as '<length>\\n<json>'")),
        (("--frontend",), dict(default="fp32", choices=["fp32", "tc"], help="resampler arithmetic: canonical float32 (default) or exact integers on the tensor cores (rates 2 ... 32 x 1.92 MHz)")),
        (("--full-scale",), dict(type=float, default=0.0, metavar="A", help="with --frontend tc on fc32 input: the range of the source (samples are taken as 23-bit fixed point over +-A)")),
    ]
    for flags, kw in options:
        ap.add_argument(*flags, **kw)
    return ap.parse_args(argv)


if __name__ == "__main__":
    main(parse())
