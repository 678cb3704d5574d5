"""The oracle against what was TRANSMITTED, not against itself: an anchor that does not share a line with
oracle/ltetrigger_oracle.c.  synth.py builds an LTE downlink from 36.211 (Zadoff-Chu PSS, m-sequence SSS, CRS, PBCH)
with a known cell id, cyclic-prefix type, frame timing and carrier offset; the restated search must report exactly
those: cell_id and cp_type on every tagged half-frame, the emitted half-frame starting on the transmitted subframe-0/5
boundary (to the sample at the search rate; shifted by the decimator's group delay (ntaps - 1) / 2 input samples when
there is one), and mean_cfo within 0.01 (150 Hz: it is a mean that includes the first noisy estimates) of the
offset in units of the 15 kHz subcarrier spacing
(srslte_pss_cfo_compute's unit, lib/pss_impl.cc:197-199).  The GPU path is bit-identical to the oracle
(tests/test_gpu_*.py), so this pins both to the physical answer."""
import numpy as np
import pytest

F_TRACKING, F_CELL, F_CP_NORM = 0x08, 0x40, 0x80      # include/ltetrigger_b200.h LTB_F_*


def _case(oracle, cell, decim, offset, cfo_hz, ext_cp, snr_db, seed, n_frames=30):
    from ltetrigger_b200 import synth
    x = synth.capture(cell, 19200 * decim * n_frames, snr_db=snr_db, decim=decim, seed=seed, offset=offset,
                      cfo_hz=cfo_hz, ext_cp=ext_cp)
    recs = oracle.trigger_run(x[None, :], decim=decim, psr_threshold=4.0, conv_mode=oracle.CONV_OS)
    return recs[(recs["flags"] & F_CELL) != 0]


def test_search_rate_captures_report_the_transmitted_cell_timing_and_offset(oracle):
    rng = np.random.default_rng(20260718)
    for i in range(16):
        cell = int(rng.integers(0, 504))
        offset = int(rng.integers(0, 19200))
        cfo_hz = float(rng.uniform(-2500.0, 2500.0)) if i % 4 else 0.0
        ext_cp = bool(i % 5 == 3)
        t = _case(oracle, cell, 1, offset, cfo_hz, ext_cp, snr_db=float(rng.uniform(6.0, 15.0)), seed=100 + i)
        what = (i, cell, offset, cfo_hz, ext_cp)
        assert len(t) >= 4, what
        assert set(t["cell_id"].tolist()) == {cell}, what
        assert set(((t["flags"] & F_CP_NORM) != 0).tolist()) == {not ext_cp}, what
        assert (t["n_id_2"] == cell % 3).all() and (t["n_id_1"] == cell // 3).all(), what
        # aligned half-frames start where the transmitter put subframe 0 or 5
        assert set((t["emit_start"] % 9600).tolist()) == {(-offset) % 9600}, what
        trk = t[(t["flags"] & F_TRACKING) != 0]
        assert len(trk) and abs(float(trk["mean_cfo"][-1]) - cfo_hz / 15000.0) < 0.01, (what, trk["mean_cfo"][-1])


@pytest.mark.parametrize("decim,ntaps", [(4, 131), (16, 525)])
def test_decimated_captures_report_the_transmitted_timing_after_the_group_delay(oracle, decim, ntaps):
    """Even cases put the delayed frame boundary on a whole search-rate sample (timing exact, carrier offset checked);
    odd cases leave it fractional: timing to the nearest sample, and no statement about the carrier offset, because a
    Zadoff-Chu sequence sampled a fraction of a sample off looks frequency-shifted to the half-symbol estimator."""
    rng = np.random.default_rng(decim)
    half = (ntaps - 1) // 2                      # group delay in input samples
    for i in range(4):
        cell = int(rng.integers(0, 504))
        offset = int(rng.integers(0, 19200 * decim))
        if i % 2 == 0:
            offset += (half - offset) % decim
        cfo_hz = float(rng.uniform(-2500.0, 2500.0))
        t = _case(oracle, cell, decim, offset, cfo_hz, False, snr_db=12.0, seed=7 + i, n_frames=24)
        what = (decim, cell, offset, cfo_hz)
        assert len(t) >= 4 and set(t["cell_id"].tolist()) == {cell}, what
        truth = (-offset / decim + half / decim) % 9600.0
        d = (t["emit_start"] % 9600 - truth + 4800.0) % 9600.0 - 4800.0
        assert np.abs(d).max() <= (0.0 if i % 2 == 0 else 1.0), (what, truth, set((t["emit_start"] % 9600).tolist()))
        if i % 2 == 0:
            trk = t[(t["flags"] & F_TRACKING) != 0]
            assert len(trk) and abs(float(trk["mean_cfo"][-1]) - cfo_hz / 15000.0) < 0.01, what
