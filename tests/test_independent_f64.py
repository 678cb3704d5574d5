"""The oracle against a second, independent statement of the same path (tests/independent_f64.py: numpy float64,
np.convolve / np.fft, written from SURVEY Appendix A and lib/pss_impl.cc / lib/sss_impl.cc, no code shared with
oracle/).  The north_star's bar, applied to the pair: window starts, threshold decisions, tracking state, peak index,
emitted half-frame start, CP type, m0 / m1, N_id_1 and cell_id identical; PSR and correlation peak value within
1e-4 relative; carrier-offset estimates within 1e-4 subcarrier spacings.  The SSS correlation values get 1e-3: on
tracking half-frames they are taken after srslte_cfo_correct, whose phasor is a 4096-entry table indexed by a
float32 phase accumulator (Appendix A.3; oracle cfo_correct) -- its rounding drift moves the index by one entry
(1.5 mrad) on part of the symbol, which is the reference's defined behaviour and not an evaluation error; the
float64 chain accumulates the phase exactly.  The GPU path is bit-identical to the oracle
(tests/test_gpu_*.py), so this bounds the distance of both from a double-precision evaluation of the reference's
algorithm."""
import numpy as np
import pytest

from conftest import load_fixture
import independent_f64 as F

F_SEARCHED, F_OVER, F_EMIT, F_TRACKING, F_TAG_LOST, F_SSS, F_CELL, F_CP_NORM = 1, 2, 4, 8, 0x10, 0x20, 0x40, 0x80
RTOL = 1e-4
SSS_RTOL = 1e-3


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    r[both_nan] = 0.0
    r[(a == 0) & (b == 0)] = 0.0
    return float(np.max(r)) if r.size else 0.0


def _compare(oracle, y32, y64, thr=4.0, conv_modes=None, what=""):
    """y32: the search-rate stream the oracle sees (complex64); y64: the same stream for the float64 chain."""
    stats = dict(windows=0, emitted=0, cells=0, psr=0.0, peak=0.0, sss=0.0, cfo=0.0)
    for conv in conv_modes or (oracle.CONV_DIRECT, oracle.CONV_FFT, oracle.CONV_OS):
        for r in range(3):
            got = oracle.chain_run(y32, r, psr_threshold=thr, conv_mode=conv)
            want = F.Chain(r, thr=thr).run(y64)
            tag = (what, conv, r)
            assert len(got) == len(want) and len(got) > 0, (tag, len(got), len(want))
            fl = got["flags"]
            for name, bit in (("searched", F_SEARCHED), ("over", F_OVER), ("emit", F_EMIT), ("tracking", F_TRACKING),
                              ("tag_lost", F_TAG_LOST), ("sss", F_SSS)):
                assert ((fl & bit) != 0).tolist() == [w[name] for w in want], (tag, name)
            for name in ("win_start", "peak_pos", "score", "emit_start", "m0", "m1", "n_id_1", "cell_id"):
                assert got[name].tolist() == [w[name] for w in want], (tag, name)
            sss = (fl & F_SSS) != 0
            assert ((fl & F_CP_NORM) != 0)[sss].tolist() == [w["cp_norm"] for w in want if w["sss"]], (tag, "cp")
            stats["psr"] = max(stats["psr"], _rel(got["psr"], [w["psr"] for w in want]))
            stats["peak"] = max(stats["peak"], _rel(got["peak_value"], [w["peak_value"] for w in want]))
            if sss.any():
                stats["sss"] = max(stats["sss"], _rel(got["m0_val"][sss], [w["m0_val"] for w in want if w["sss"]]),
                                   _rel(got["m1_val"][sss], [w["m1_val"] for w in want if w["sss"]]))
            for name in ("cfo", "mean_cfo"):
                d = np.abs(got[name].astype(np.float64) - np.array([w[name] for w in want]))
                stats["cfo"] = max(stats["cfo"], float(d.max()))
            stats["windows"] += len(got)
            stats["emitted"] += int(((fl & F_EMIT) != 0).sum())
            stats["cells"] += int(((fl & F_CELL) != 0).sum())
    assert stats["psr"] < RTOL and stats["peak"] < RTOL and stats["sss"] < SSS_RTOL and stats["cfo"] < 1e-4, (what, stats)
    print(what, stats)
    return stats


def test_pss_filter_and_decimator_taps_agree(oracle):
    for r in range(3):
        assert np.max(np.abs(oracle.pss_taps(r) - F.pss_filter(r))) < 2e-8        # float32 rounding of |h| <= 0.02
    for d, n in ((4, 131), (8, 263), (16, 525), (12, 393)):
        t = F.decimator_taps(d)
        assert len(t) == n == len(oracle.decim_taps(d))
        assert np.max(np.abs(oracle.decim_taps(d) - t)) < 1e-8
    c0, c1, s, z, tab = oracle.sss_tables(2)
    i = np.arange(31)
    assert (s == F.S_T).all() and (z == F.Z_T).all() and (c0 == F.C_T[(i + 2) % 31]).all()
    assert (tab == F.N_ID_1[:30, 1:]).all()


def test_6prb_fixture_all_three_evaluations(oracle):
    x, decim, cell = load_fixture("6prb", seconds=0.4)
    s = _compare(oracle, x, x, what="6prb")
    assert s["cells"] > 0 and s["emitted"] > 50, s


@pytest.mark.parametrize("name", ["25prb", "100prb"])
def test_decimated_fixtures(oracle, name):
    """The decimator too is evaluated independently (float64 taps from the design formulas, one matrix product)."""
    x, decim, cell = load_fixture(name, seconds=0.25)
    y32 = oracle.decimate(x, decim)
    y64 = F.decimate(x, decim)
    assert len(y32) == len(y64)
    assert np.max(np.abs(y32 - y64)) < 1e-6 * np.max(np.abs(y64))
    s = _compare(oracle, y32, y64, conv_modes=(oracle.CONV_OS,), what=name)
    assert s["cells"] > 0, s


def test_synthetic_captures(oracle):
    """Seeded synthetic cells: noise, carrier offset, extended CP, random timing; and one noise-only stream (no decision
    may differ there either: 3 x 40 threshold comparisons on pure noise)."""
    from ltetrigger_b200 import synth
    rng = np.random.default_rng(515)
    tot = dict(windows=0, cells=0)
    for i in range(6):
        cell = int(rng.integers(0, 504))
        x = synth.capture(cell, 19200 * 14, snr_db=float(rng.uniform(0.0, 10.0)), seed=40 + i,
                          offset=int(rng.integers(0, 19200)), cfo_hz=float(rng.uniform(-3000, 3000)), ext_cp=(i == 3))
        s = _compare(oracle, x, x, conv_modes=(oracle.CONV_OS, oracle.CONV_DIRECT), what="synth %d cell %d" % (i, cell))
        tot["windows"] += s["windows"]
        tot["cells"] += s["cells"]
    x = synth.capture(0, 19200 * 10, snr_db=0.0, seed=3, noise_only=True)
    s = _compare(oracle, x, x, conv_modes=(oracle.CONV_OS,), what="noise only")
    assert s["emitted"] == 0 and s["psr"] > 0.0, s
    x = synth.capture(0, 19200 * 4, seed=3, noise_only=True)                 # silence: PSR is 0 / 0 on both sides
    _compare(oracle, x, x, conv_modes=(oracle.CONV_OS, oracle.CONV_DIRECT), what="silence")
    assert tot["cells"] > 20, tot


def test_integer_front_end_against_float64(oracle):
    """LTB_FRONTEND_TC_INT's arithmetic (int8 digit products, 4-digit taps; fc32 taken as 23-bit fixed point) against the
    float64 decimator on the same samples, then the whole chain on its output: identical decisions."""
    from ltetrigger_b200 import synth
    x, decim, cell = load_fixture("100prb", seconds=0.25)
    iq = synth.to_sc16(x[None, :])[0]
    y_int = oracle.decimate_tcint_sc16(iq, decim)
    y64 = F.decimate(oracle.sc16_to_fc32(iq).astype(np.complex128), decim)
    assert np.max(np.abs(y_int - y64)) < 1e-6 * np.max(np.abs(y64))
    s = _compare(oracle, y_int, y64, conv_modes=(oracle.CONV_OS,), what="100prb sc16, integer front end")
    assert s["cells"] > 0, s
    fs = float(8 * np.sqrt(np.mean(np.abs(x) ** 2)))
    y_fix = oracle.decimate_tcint_fc32(x, fs, decim)
    y64 = F.decimate(x, decim)
    assert np.max(np.abs(y_fix - y64)) < 2e-6 * np.max(np.abs(y64))
    s = _compare(oracle, y_fix, y64, conv_modes=(oracle.CONV_OS,), what="100prb fc32 as fixed point, integer front end")
    assert s["cells"] > 0, s


def test_threshold_churn(oracle):
    """Threshold 1.7 at low SNR: chains gain and lose tracking repeatedly (reset_score, tracking_lost tag, EMA reset,
    forced emission after a loss) -- every transition identical in both statements."""
    from ltetrigger_b200 import synth
    lost = 0
    for i, snr in enumerate((-6.0, -3.0, -8.0)):
        x = synth.capture(77 + 100 * i, 19200 * 20, snr_db=snr, seed=900 + i, cfo_hz=1200.0 * (i - 1))
        for r in range(3):
            got = oracle.chain_run(x, r, psr_threshold=1.7, conv_mode=oracle.CONV_OS)
            lost += int(((got["flags"] & F_TAG_LOST) != 0).sum())
        _compare(oracle, x, x, thr=1.7, conv_modes=(oracle.CONV_OS,), what="churn %g dB" % snr)
    assert lost > 10, lost
