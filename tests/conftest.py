import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gr-ltetrigger_b200", "python"))

GOLDEN = os.path.join(ROOT, "tests", "golden")
FIXTURES = {          # name: (file, decim, cell_id)  python/qa_downlink_trigger_c.py:67-203
    "6prb": ("lte_frame_6prb_cellid_123", 1, 123),
    "25prb": ("lte_frame_25prb_cellid_124", 4, 124),
    "50prb": ("lte_frame_50prb_cellid_125", 8, 125),
    "100prb": ("lte_frame_100prb_cellid_369", 16, 369),
}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_fixture(name, seconds=1.0):
    """file_source(repeat=True) -> head(seconds) as in the reference's QA flowgraphs."""
    fname, decim, cell_id = FIXTURES[name]
    x = np.fromfile(os.path.join(GOLDEN, "test_frames", fname), np.complex64)
    n = int(round(seconds * 1.92e6)) * decim
    n -= n % (8 * decim)
    reps = -(-n // len(x))
    return np.tile(x, reps)[:n], decim, cell_id


def has_gpu():
    try:
        import ltetrigger_b200 as lt
        return lt.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


def assert_recs_equal(got, want, float_fields_exact=True):
    """Field-by-field comparison of WINDOW_REC arrays; floats compared as bit patterns."""
    assert len(got) == len(want), "record count %d != %d" % (len(got), len(want))
    for name in want.dtype.names:
        g, w = got[name], want[name]
        if g.dtype.kind == "f":
            gb, wb = g.view(np.uint32), w.view(np.uint32)
            same = (gb == wb) | ((g == 0) & (w == 0)) | (np.isnan(g) & np.isnan(w))   # +0 == -0; NaN payloads (0/0 on
            # silent input: x86 gives the negative quiet NaN, the GPU the canonical one) are not part of the contract
            if not same.all():
                i = int(np.argmin(same))
                raise AssertionError("field %s differs at record %d: got %r want %r (stream %d root %d win %d)"
                                     % (name, i, g[i], w[i], want["stream"][i], want["n_id_2"][i], want["win_index"][i]))
        else:
            if not (g == w).all():
                i = int(np.argmin(g == w))
                raise AssertionError("field %s differs at record %d: got %r want %r (stream %d root %d win %d)"
                                     % (name, i, g[i], w[i], want["stream"][i], want["n_id_2"][i], want["win_index"][i]))


def snr_demo_capture(seconds, seed, snr_db=-10.405):
    """The state the reference documents in docs/gr_ltetrigger_snr_demo.png (examples/snr_ltetrigger.grc):
    the 6 PRB test frame plus Gaussian noise at a measured SNR of -10.405 dB (signal power over noise
    power across the 1.92 MHz band), demod threshold 1.7 -> "Tracking: True, Cell ID: 123, PRBs: 6,
    PHICH Resources: '1', CP Mode: 'Normal'"."""
    x, _, _ = load_fixture("6prb", seconds)
    rng = np.random.default_rng(seed)
    pn = np.mean(np.abs(x) ** 2) / 10 ** (snr_db / 10)
    noise = np.sqrt(pn / 2) * (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x)))
    return (x + noise).astype(np.complex64)
