"""Seeded campaign of the oracle against the independent float64 statement (tests/independent_f64.py): random cells,
SNR, timing, carrier offset, CP type, threshold, decimation (through both decimators) and matched-filter evaluation;
counts windows compared and differences by kind.  A decision that differs is not necessarily an error of either side -- a
float32 and a float64 evaluation may fall on different sides of a tie -- so differences are reported with the margin that
decided them, not asserted away.
  python tests/independent_campaign.py --seconds 300 --seed 1 > profiles/independent_f64_campaign_rNN.txt"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gr-ltetrigger_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

F_BITS = (("searched", 1), ("over", 2), ("emit", 4), ("tracking", 8), ("tag_lost", 0x10), ("sss", 0x20))


def one_case(rng, O, F, synth):
    cell = int(rng.integers(0, 504))
    decim = int(rng.choice([1, 1, 4, 8, 16]))
    snr = float(rng.uniform(-8.0, 12.0))
    thr = float(rng.choice([4.0, 4.0, 2.5, 1.7]))
    ext = bool(rng.random() < 0.2)
    cfo = float(rng.uniform(-4000.0, 4000.0)) if rng.random() < 0.7 else 0.0
    frames = int(rng.integers(12, 32))
    conv = int(rng.choice([O.CONV_OS, O.CONV_OS, O.CONV_DIRECT, O.CONV_FFT]))
    x = synth.capture(cell, 19200 * decim * frames, snr_db=snr, decim=decim, seed=int(rng.integers(1 << 30)),
                      cfo_hz=cfo, ext_cp=ext, noise_only=bool(rng.random() < 0.1))
    fe = str(rng.choice(["fp32", "fp32", "tc-sc16", "tc-fc32"])) if decim > 1 else "fp32"
    if fe == "tc-sc16":                                       # LTB_FRONTEND_TC_INT restated, on int16 wire samples
        iq = synth.to_sc16(x[None, :])[0]
        y32, y64 = O.decimate_tcint_sc16(iq, decim), F.decimate(O.sc16_to_fc32(iq).astype(np.complex128), decim)
    elif fe == "tc-fc32":                                     # the same on fc32 taken as 23-bit fixed point over 8 rms
        y32, y64 = O.decimate_tcint_fc32(x, float(8 * np.sqrt(np.mean(np.abs(x) ** 2))), decim), F.decimate(x, decim)
    else:
        y32, y64 = (O.decimate(x, decim) if decim > 1 else x), F.decimate(x, decim)
    label = "cell=%d D=%d fe=%s snr=%.1f thr=%.1f ext=%d cfo=%.0f frames=%d conv=%d" % (cell, decim, fe, snr, thr, ext, cfo, frames, conv)
    res = dict(windows=0, differ=[], psr=0.0, peak=0.0, cfo=0.0, sss=0.0, cells=0)
    for r in range(3):
        got = O.chain_run(y32, r, psr_threshold=thr, conv_mode=conv)
        want = F.Chain(r, thr=thr).run(y64)
        n = min(len(got), len(want))
        first = None
        for i in range(n):
            g, w = got[i], want[i]
            same = all(bool(g["flags"] & b) == w[k] for k, b in F_BITS) and \
                all(int(g[k]) == w[k] for k in ("win_start", "peak_pos", "score", "emit_start", "m0", "m1", "n_id_1", "cell_id"))
            if not same:
                first = i
                break
        upto = n if first is None else first
        res["windows"] += upto
        if first is not None or len(got) != len(want):
            i = upto
            if i < n:
                g, w = got[i], want[i]
                what = [k for k, b in F_BITS if bool(g["flags"] & b) != w[k]] + \
                       [k for k in ("win_start", "peak_pos", "score", "emit_start", "m0", "m1", "n_id_1", "cell_id") if int(g[k]) != w[k]]
                res["differ"].append("root %d window %d: %s; psr %.7g / %.7g (thr %.2f), peak_pos %d / %d, m0 m1 %d %d / %d %d"
                                     % (r, i, ",".join(what), g["psr"], w["psr"], thr, g["peak_pos"], w["peak_pos"],
                                        g["m0"], g["m1"], w["m0"], w["m1"]))
            else:
                res["differ"].append("root %d: %d / %d windows" % (r, len(got), len(want)))
        if upto:
            g = got[:upto]
            rel = lambda a, b: float(np.nanmax(np.abs(a.astype(np.float64) - b) / np.maximum(np.abs(b), 1e-300))) if len(a) else 0.0  # noqa: E731
            ok = ~np.isnan(g["psr"])
            res["psr"] = max(res["psr"], rel(g["psr"][ok], np.array([w["psr"] for w in want[:upto]])[ok]))
            res["peak"] = max(res["peak"], rel(g["peak_value"], np.array([w["peak_value"] for w in want[:upto]])))
            res["cfo"] = max(res["cfo"], float(np.max(np.abs(g["mean_cfo"].astype(np.float64) - np.array([w["mean_cfo"] for w in want[:upto]])))))
            s = (g["flags"] & 0x20) != 0
            if s.any():
                res["sss"] = max(res["sss"], rel(g["m0_val"][s], np.array([w["m0_val"] for w in want[:upto] if w["sss"]])))
            res["cells"] += int(((g["flags"] & 0x40) != 0).sum())
    return label, res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    from oracle import oracle as O
    from ltetrigger_b200 import synth
    import independent_f64 as F
    rng = np.random.default_rng(a.seed)
    t0, n, tot = time.time(), 0, dict(windows=0, cells=0, differ=0, psr=0.0, peak=0.0, cfo=0.0, sss=0.0)
    while time.time() - t0 < a.seconds:
        label, res = one_case(rng, O, F, synth)
        n += 1
        tot["windows"] += res["windows"]
        tot["cells"] += res["cells"]
        tot["differ"] += len(res["differ"])
        for k in ("psr", "peak", "cfo", "sss"):
            tot[k] = max(tot[k], res[k])
        print("%s %4d %s windows=%d cells=%d psr=%.2e peak=%.2e cfo=%.2e sss=%.2e" %
              ("ok  " if not res["differ"] else "DIFF", n, label, res["windows"], res["cells"], res["psr"], res["peak"], res["cfo"], res["sss"]))
        for d in res["differ"]:
            print("       " + d)
        sys.stdout.flush()
    print("independent_campaign: seed %d, %d cases, %d windows compared up to the first difference of a chain, %d cell-tagged, "
          "%d chains with a difference; max rel diff PSR %.2e, peak %.2e, SSS value %.2e; max abs diff mean_cfo %.2e"
          % (a.seed, n, tot["windows"], tot["cells"], tot["differ"], tot["psr"], tot["peak"], tot["sss"], tot["cfo"]))


if __name__ == "__main__":
    main()
