"""Host-side MIB decode (ltb_mib_decode + the `mib` block mirror) against the values the
reference's own tests pin (python/qa_downlink_trigger_c.py:46-65): nof_prb, phich_len,
nof_phich_resources, nof_tx_ports, cp_len, cell_id for the four bundled test_frames.  Runs
without a GPU: the oracle's restated pss/sss blocks produce the tagged, aligned, CFO-corrected
half-frames, the product's host code decodes them."""
import numpy as np
import pytest

from conftest import FIXTURES, load_fixture, snr_demo_capture

NOF_PRB = {"6prb": 6, "25prb": 25, "50prb": 50, "100prb": 100}


@pytest.mark.parametrize("name", list(FIXTURES))
def test_mib_block_on_oracle_chain(oracle, name):
    import ltetrigger_b200 as lt
    x, decim, cell_id = load_fixture(name, 0.25)
    y = oracle.decimate(x, decim) if decim > 1 else x
    k = cell_id % 3
    op, os_, mb = oracle.Pss(k, 4.0), oracle.Sss(k), lt.mib(exit_on_success=True)
    tracked = []
    mb.msg_connect("track", tracked.append)
    buf = np.concatenate([np.zeros(960, np.complex64), y])
    pos, written, n_tried = 960, 0, 0
    while pos - 960 + oracle.LOOKAHEAD <= len(y) and not mb.done:
        nout, ncons, out, rec = op.work(buf, pos)
        if nout:
            lost = bool(rec["flags"] & oracle.F_TAG_LOST)
            _, srec = os_.work(out, lost)
            tags = [lt.tag_t(written, "tracking_lost", None)] if lost else []
            if srec["flags"] & oracle.F_CELL:
                tags += [lt.tag_t(written, "cell_id", int(srec["cell_id"])),
                         lt.tag_t(written, "cp_type", bool(srec["flags"] & oracle.F_CP_NORM))]
                n_tried += 1
            mb._in_tags, mb._nitems_read = tags, written
            mb.general_work(9600, [9600], [out], [None])
            written += nout
        pos += ncons
    assert mb.done and len(tracked) == 1 and n_tried <= 2      # SF0 or SF5 first: found within two half-frames
    cell = tracked[0]
    assert cell["cell_id"] == cell_id and cell["cp_len"] == "Normal"
    assert cell["nof_prb"] == NOF_PRB[name] and cell["nof_tx_ports"] == 1
    assert cell["phich_len"] == "Normal" and cell["nof_phich_resources"] == "1"
    assert set(cell) == {"cell_id", "nof_tx_ports", "cp_len", "nof_prb", "phich_len", "nof_phich_resources",
                         "sfn_offset", "tracking_start_time"}      # lib/mib_impl.cc:185-251


@pytest.mark.parametrize("seed", [0, 1])
def test_reference_snr_demo_screenshot(oracle, seed):
    """docs/gr_ltetrigger_snr_demo.png of the reference: at -10.4 dB and threshold 1.7 the trigger is
    tracking cell 123 and has decoded 6 PRB, PHICH resources '1', normal CP.  Same state here (oracle
    chain N_id_2 = 0 + host mib), within 8 s of signal; junk cell ids from the SSS at this SNR never
    pass the PBCH CRC."""
    import ltetrigger_b200 as lt
    y = snr_demo_capture(8.0, seed)
    op, os_, mb = oracle.Pss(0, 1.7), oracle.Sss(0), lt.mib(exit_on_success=True)
    tracked = []
    mb.msg_connect("track", tracked.append)
    buf = np.concatenate([np.zeros(960, np.complex64), y])
    pos, written = 960, 0
    while pos - 960 + oracle.LOOKAHEAD <= len(y) and not mb.done:
        nout, ncons, out, rec = op.work(buf, pos)
        if nout:
            lost = bool(rec["flags"] & oracle.F_TAG_LOST)
            _, srec = os_.work(out, lost)
            tags = [lt.tag_t(written, "tracking_lost", None)] if lost else []
            if srec["flags"] & oracle.F_CELL:
                tags += [lt.tag_t(written, "cell_id", int(srec["cell_id"])),
                         lt.tag_t(written, "cp_type", bool(srec["flags"] & oracle.F_CP_NORM))]
            mb._in_tags, mb._nitems_read = tags, written
            mb.general_work(9600, [9600], [out], [None])
            written += nout
        pos += ncons
    assert mb.done and len(tracked) == 1
    cell = tracked[0]
    assert (cell["cell_id"], cell["nof_prb"], cell["nof_phich_resources"], cell["cp_len"]) == (123, 6, "1", "Normal")
    assert op.tracking_score() > 0                                       # "Tracking: True"


@pytest.mark.parametrize("n_ports,nof_prb,decim,ext_cp", [(1, 15, 2, False), (2, 6, 1, False), (2, 75, 12, False),
                                                         (2, 25, 4, True), (4, 50, 8, False), (4, 6, 1, True)])
def test_mib_on_synthetic_cells_one_two_and_four_ports(oracle, n_ports, nof_prb, decim, ext_cp):
    """Synthetic cells with CRS and PBCH (synth.py's transmitter: CRC mask, tail-biting code, rate
    matching, scrambling, transmit diversity: SFBC for two ports, SFBC-FSTD for four) through the oracle chain and
    the host mib: bandwidths the bundled frames do not cover (15 and 75 PRB), extended CP, and two / four antenna
    ports -- what srslte_ue_mib_decode reports as nof_ports (lib/mib_impl.cc:163-166)."""
    import ltetrigger_b200 as lt
    from ltetrigger_b200 import synth
    cell_id = 100 + 3 * nof_prb + n_ports
    x = synth.capture(cell_id, 19200 * decim * 14, snr_db=12.0, decim=decim, seed=nof_prb, ext_cp=ext_cp,
                      mib=dict(nof_prb=nof_prb, n_ports=n_ports, phich_res=1, sfn0=4,
                               h=(0.9 + 0.2j, -0.4 + 0.7j, 0.3 - 0.8j, -0.6 - 0.5j)))
    y = oracle.decimate(x, decim) if decim > 1 else x
    k = cell_id % 3
    op, os_, mb = oracle.Pss(k, 4.0), oracle.Sss(k), lt.mib(exit_on_success=True)
    tracked = []
    mb.msg_connect("track", tracked.append)
    buf = np.concatenate([np.zeros(960, np.complex64), y])
    pos, written = 960, 0
    while pos - 960 + oracle.LOOKAHEAD <= len(y) and not mb.done:
        nout, ncons, out, rec = op.work(buf, pos)
        if nout:
            lost = bool(rec["flags"] & oracle.F_TAG_LOST)
            _, srec = os_.work(out, lost)
            tags = [lt.tag_t(written, "tracking_lost", None)] if lost else []
            if srec["flags"] & oracle.F_CELL:
                tags += [lt.tag_t(written, "cell_id", int(srec["cell_id"])),
                         lt.tag_t(written, "cp_type", bool(srec["flags"] & oracle.F_CP_NORM))]
            mb._in_tags, mb._nitems_read = tags, written
            mb.general_work(9600, [9600], [out], [None])
            written += nout
        pos += ncons
    assert mb.done and len(tracked) == 1
    cell = tracked[0]
    assert (cell["cell_id"], cell["nof_prb"], cell["nof_tx_ports"]) == (cell_id, nof_prb, n_ports)
    assert cell["cp_len"] == ("Extended" if ext_cp else "Normal") and cell["nof_phich_resources"] == "1/2"
    # "sfn_offset" is what the reference publishes under that key: the MIB's 8-bit SFN field << 2
    # (lib/mib_impl.cc:167-172 hands &d_sfn_offset to srslte_pbch_mib_unpack as its sfn output).  The
    # capture tiles four frames with SFN 4..7 (synth.capture), so the field is 1 and the value 4.
    assert cell["sfn_offset"] == 4, cell["sfn_offset"]


def test_mib_decode_rejects_noise_sf5_and_bad_arguments():
    import ctypes as C
    import ltetrigger_b200 as lt
    from ltetrigger_b200 import _abi as A
    L, m = lt.lib(), A.Mib()
    rng = np.random.default_rng(1)
    noise = (rng.standard_normal(9600) + 1j * rng.standard_normal(9600)).astype(np.complex64)
    assert L.ltb_mib_decode(noise.ctypes.data, 123, 1, C.byref(m)) == 0
    zeros = np.zeros(9600, np.complex64)
    assert L.ltb_mib_decode(zeros.ctypes.data, 123, 1, C.byref(m)) == 0
    x, _, cell_id = load_fixture("6prb", 0.02)
    assert L.ltb_mib_decode(x[:9600].ctypes.data, cell_id, 1, C.byref(m)) == 1 and m.nof_prb == 6
    assert L.ltb_mib_decode(x[9600:19200].ctypes.data, cell_id, 1, C.byref(m)) == 0      # subframe 5: no PBCH
    assert L.ltb_mib_decode(x[:9600].ctypes.data, cell_id + 1, 1, C.byref(m)) == 0        # wrong cell: CRS/scrambling differ
    assert L.ltb_mib_decode(x[:9600].ctypes.data, 504, 1, C.byref(m)) == lt.ERROR_INVALID_INPUTS
    assert L.ltb_mib_decode(None, 1, 1, C.byref(m)) == lt.ERROR_INVALID_INPUTS


def test_mib_block_drop_protocol():
    """lib/mib_impl.cc:107-125: tracking_lost -> drop with the identical object, then silence until
    the next decode; untagged or doubly tagged half-frames are ignored."""
    import ltetrigger_b200 as lt
    x, _, cell_id = load_fixture("6prb", 0.02)
    hf = x[:9600]
    mb = lt.mib()
    tracked, dropped = [], []
    mb.msg_connect("track", tracked.append)
    mb.msg_connect("drop", dropped.append)

    def feed(tags):
        mb._in_tags, mb._nitems_read = tags, 0
        return mb.general_work(9600, [9600], [hf], [None])

    good = [lt.tag_t(0, "cell_id", cell_id), lt.tag_t(0, "cp_type", True)]
    assert feed([]) == 0 and not tracked                                   # no tags: swallowed
    assert feed(good + [lt.tag_t(0, "cell_id", 5)]) == 0 and not tracked   # two cell_id tags: sanity check trips
    assert feed(good) == 9600 and len(tracked) == 1
    assert feed(good) == 0 and len(tracked) == 1                           # already published
    assert feed([lt.tag_t(0, "tracking_lost", None)]) == 0 and dropped == tracked and dropped[0] is tracked[0]
    assert feed([lt.tag_t(0, "tracking_lost", None)]) == 0 and len(dropped) == 1   # nothing published: no second drop
    assert feed(good) == 9600 and len(tracked) == 2


def test_cellstore_track_drop_by_identity():
    """lib/cellstore_impl.cc:46-105 (python/qa_cellstore.py is an empty shell in the reference): "track"
    appends, "drop" removes the identical object only, accessors as include/ltetrigger/cellstore.h:59-65."""
    import ltetrigger_b200 as lt
    store = lt.cellstore()
    assert store.message_ports() == ["track", "drop"] and not store.tracking() and store.latest_cell() is None
    a, b, twin = {"cell_id": 1}, {"cell_id": 2}, {"cell_id": 1}
    store.track_cell(a)
    store.track_cell(b)
    assert store.tracking() and store.cells() == [a, b] and store.latest_cell() is b
    store.drop_cell(twin)                          # equal but not the same object: stays (pmt identity, :103)
    assert store.cells() == [a, b]
    store.drop_cell(a)
    assert store.cells() == [b] and store.latest_cell() is b
    store.drop_cell(b)
    assert not store.tracking() and store.cells() == []
    # wired to a mib stage like examples/cell_search_file.py:82-88
    mb = lt.mib()
    store.connect(mb)
    mb._pub("track", a)
    assert store.latest_cell() is a
    mb._pub("drop", a)
    assert not store.tracking()
