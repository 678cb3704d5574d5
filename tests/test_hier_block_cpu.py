"""The hier-block mirror `downlink_trigger_c` (python/downlink_trigger_c.py:13-73), its three host `mib` blocks and
`cellstore` on CPU, with the oracle's restated blocks standing in for the CUDA engine behind `Trigger`'s interface: the
reference's QA flowgraphs (python/qa_downlink_trigger_c.py:67-203) and the host logic around the engine -- message ports,
threshold clamp, tag offsets per chain, records, exit_on_success -- without a device.  tests/test_gpu_blocks.py runs the
same assertions on the real engine."""
import numpy as np
import pytest

from conftest import FIXTURES, load_fixture

NOF_PRB = {"6prb": 6, "25prb": 25, "50prb": 50, "100prb": 100}


def oracle_engine(oracle):
    """A class with Trigger's constructor and the methods the block mirrors call, computing with oracle.Pss / oracle.Sss
    call by call under the engine's scheduler rule (a window runs once 18365 samples from its start have arrived)."""
    import ltetrigger_b200 as lt
    from ltetrigger_b200 import _abi as A

    class Engine:
        def __init__(self, n_streams, decim=1, psr_threshold=4.0, max_chunk=1 << 18, root_mask=7, track_after=16,
                     track_every=8, **kw):
            assert n_streams == 1
            self.max_chunk, self.thr, self.decim = max_chunk, psr_threshold, decim
            self.roots = [k for k in range(3) if root_mask >> k & 1]
            self.raw, self.n_y = np.zeros(0, np.complex64), 0    # decim > 1: the fused resampler, restated by orc_decimate
            self.pss = [oracle.Pss(k, psr_threshold, track_after, track_every, conv_mode=oracle.CONV_DIRECT) for k in range(3)]
            self.sss = [oracle.Sss(k) for k in range(3)]
            self.buf = np.zeros(960, np.complex64)               # the zero history GNU Radio puts in front
            self.pos = [960, 960, 960]
            self.hfs = []

        def process(self, x):
            x = np.asarray(x[0], np.complex64)
            if self.decim > 1:                                   # y[k] depends on past input only: decimate all, keep the new outputs
                self.raw = np.concatenate([self.raw, x])
                y = oracle.decimate(self.raw, self.decim)
                x, self.n_y = y[self.n_y:], len(y)
            self.buf = np.concatenate([self.buf, x])
            out = []
            for k in self.roots:
                while self.pos[k] - 960 + oracle.LOOKAHEAD <= len(self.buf) - 960:
                    nout, ncons, hf, rec = self.pss[k].work(self.buf, self.pos[k])
                    rec = rec.copy()
                    rec["win_start"] = self.pos[k] - 960
                    rec["emit_start"] = self.pos[k] - 960 + rec["emit_start"] if nout else -1
                    if nout:
                        _, rec = self.sss[k].work(hf, bool(rec["flags"] & oracle.F_TAG_LOST), rec)
                    out.append((k, rec, hf if nout else None))
                    self.pos[k] += ncons
            recs = np.zeros(len(out), A.WINDOW_REC)
            self.hfs = []
            for i, (k, rec, hf) in enumerate(out):               # engine order: by chain, then by call
                for f in rec.dtype.names:
                    recs[i][f] = rec[f]
                if hf is not None:
                    self.hfs.append(hf)
            return recs

        def fetch_halfframes(self, n):
            assert n == len(self.hfs)
            return np.stack(self.hfs)

        def stats(self, stream, k):
            st = A.PssStats()
            p = self.pss[k]
            st.max_psr, st.mean_psr, st.mean_cfo = p.max_psr(), p.mean_psr(), p.mean_cfo()
            st.psr_threshold, st.tracking_score = p.psr_threshold(), p.tracking_score()
            return st

        def set_psr_threshold(self, t, stream=-1, n_id_2=-1, clamp=True):
            if clamp:
                t = max(t, lt.MIN_PSR_THRESHOLD)
            for k in range(3):
                if n_id_2 in (-1, k):
                    self.pss[k].set_psr_threshold(t)

    return Engine


@pytest.fixture()
def hier(oracle, monkeypatch):
    import ltetrigger_b200 as lt
    from ltetrigger_b200 import blocks
    monkeypatch.setattr(blocks, "Trigger", oracle_engine(oracle))
    return lt


@pytest.mark.parametrize("name", list(FIXTURES))
def test_reference_qa_flowgraphs(hier, oracle, name):
    lt = hier
    x, decim, cell_id = load_fixture(name, 0.3)
    y = oracle.decimate(x, decim) if decim > 1 else x            # rational_resampler_ccc(1, D) in front, as in the QA
    trig = lt.downlink_trigger_c(psr_threshold=4, exit_on_success=True)
    assert trig.message_ports() == ["track", "drop"]
    store = lt.cellstore().connect(trig)
    tracked, dropped = [], []
    trig.msg_connect("track", tracked.append)
    trig.msg_connect("drop", dropped.append)
    tags = []
    for a in range(0, len(y), 40000):                            # chunks that are no multiple of a half-frame
        tags += trig.work(y[a:a + 40000])
    assert len(tracked) == 1 and dropped == []
    cell = tracked[0]
    assert cell["cell_id"] == cell_id and cell["cp_len"] == "Normal" and cell["nof_phich_resources"] == "1"
    assert cell["nof_prb"] == NOF_PRB[name] and cell["nof_tx_ports"] == 1 and cell["phich_len"] == "Normal"
    assert store.tracking() and store.latest_cell() is cell and store.cells() == [cell]
    assert [m.done for m in (trig.mib0, trig.mib1, trig.mib2)] == [k == cell_id % 3 for k in range(3)]
    ids = [(k, t.value) for k, t in tags if t.key == "cell_id"]
    assert {v for _, v in ids} == {cell_id} and {k for k, _ in ids} == {cell_id % 3}
    assert all(t.value is True for _, t in tags if t.key == "cp_type")
    # tag offsets count the items each chain's pss has written: multiples of one half-frame, increasing per chain
    for k in range(3):
        offs = [t.offset for kk, t in tags if kk == k and t.key in ("tracking_lost", "cell_id")]
        assert all(o % 9600 == 0 for o in offs) and offs == sorted(offs)
    assert (trig.pss0, trig.pss1, trig.pss2)[cell_id % 3].tracking_score() == 16.0
    assert [p.tracking_score() for i, p in enumerate((trig.pss0, trig.pss1, trig.pss2)) if i != cell_id % 3] == [0.0, 0.0]
    # one record per general_work call of each chain; windows advance
    for k in range(3):
        ws = [int(r["win_start"]) for r in trig.records if r["n_id_2"] == k]
        assert len(ws) > 20 and ws == sorted(ws)
    # threshold clamp (python/downlink_trigger_c.py:63-73)
    trig.set_psr_threshold(0.5)
    assert trig.psr_threshold == 1.5 and trig.pss1.psr_threshold() == 1.5
    assert lt.downlink_trigger_c(psr_threshold=0.1).psr_threshold == 1.5
    trig.set_psr_threshold(7.25)
    assert trig.psr_threshold == 7.25 and trig.pss2.psr_threshold() == 7.25


def test_custom_sink_replaces_the_mib_stage(hier, oracle):
    """`mib_sink(k, tags, halfframe)`: what a flowgraph gets that connects its own block behind sss -- every emitted
    half-frame with its tags, in stream order per chain; returned (port, msg) pairs go out on the hier block's ports."""
    lt = hier
    x, _, cell_id = load_fixture("6prb", 0.2)
    trig = lt.downlink_trigger_c(psr_threshold=4)
    seen, got = [], []
    trig.msg_connect("track", got.append)

    def sink(k, tags, hf):
        seen.append((k, [t.key for t in tags], hf.shape))
        if any(t.key == "cell_id" for t in tags) and not got:
            return [("track", {"cell_id": [t.value for t in tags if t.key == "cell_id"][0]})]
    trig.mib_sink = sink
    trig.work(x)
    assert got == [{"cell_id": cell_id}]
    assert {k for k, _, _ in seen} == {cell_id % 3} and all(s == (9600,) for _, _, s in seen)
    assert seen[0][1] == ["tracking_lost"] and ["cell_id", "cp_type"] in [keys for _, keys, _ in seen]


@pytest.mark.parametrize("name,rate", [("6prb", "1.92M"), ("25prb", "7.68M")])
def test_cell_search_file_cli_like_reference_test_sh(hier, name, rate, capsys):
    """examples/test.sh of the reference (:3-6): `cell_search_file.py -s RATE FRAME --repeat --time-out 1` prints the cell as
    JSON with "status": "FOUND"; with a cut-off too short to reach tracking it prints NOT_FOUND.  (The two wider frames run
    the same code; the stand-in's decimator is too slow for them here -- they are in tests/test_gpu_blocks.py.)"""
    import importlib.util
    import json
    import os
    from conftest import GOLDEN, ROOT
    spec = importlib.util.spec_from_file_location("cell_search_file", os.path.join(ROOT, "examples", "cell_search_file.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    fname, decim, cell_id = FIXTURES[name]
    path = os.path.join(GOLDEN, "test_frames", fname)
    out = cli.main(cli.parse(["-s", rate, path, "--repeat", "--time-out", "60"]))
    cell = json.loads(out[0])
    assert cell["status"] == "FOUND" and cell["cell_id"] == cell_id and cell["nof_prb"] == NOF_PRB[name]
    assert cell["cp_len"] == "Normal" and cell["nof_tx_ports"] == 1 and len(out) == 1
    assert "Starting cell search... done." in capsys.readouterr().out
    # -c counts search-rate samples (the reference puts head() behind the resampler): 60 ms cannot reach tracking
    out = cli.main(cli.parse(["-s", rate, path, "--repeat", "-c", "115200"]))
    assert json.loads(out[0]) == {"status": "NOT_FOUND"}
    # and 200 ms do
    out = cli.main(cli.parse(["-s", rate, path, "--repeat", "-c", "384000"]))
    assert json.loads(out[0])["cell_id"] == cell_id
    with pytest.raises(SystemExit):
        cli.main(cli.parse(["-s", "2M", path]))                  # not a multiple of 1.92 MHz


@pytest.mark.parametrize("name", ["6prb", "25prb"])
def test_pss_block_mirror_scheduler_contract(hier, oracle, name):
    """The `pss` block mirror under the scheduler stand-in (gr_emu: zero-filled history, forecast, consume_each, tags) with
    the oracle behind the engine interface, against the oracle's pss block driven directly: what is under test is the
    mirror's own logic -- look-ahead queue, one window per general_work call, consume 9600 or peak - 960 + 9600, the
    "tracking_lost" tag on item 0 of half-frames emitted while not tracking, accessors (lib/pss_impl.cc:154-223)."""
    lt = hier
    from ltetrigger_b200 import gr_emu
    from ltetrigger_b200.blocks import _block, HALF_FRAME_LENGTH
    x, decim, cell_id = load_fixture(name, 0.3)
    y = oracle.decimate(x, decim) if decim > 1 else x
    k = cell_id % 3

    class PassThrough(_block):                                   # stands where sss stands; its own mirror needs the device
        def work(self, noutput_items, input_items, output_items):
            output_items[0][:HALF_FRAME_LENGTH] = input_items[0][:HALF_FRAME_LENGTH]
            return HALF_FRAME_LENGTH

    p = lt.pss(k, 4.0)
    assert p.history() == 9600 and p.output_multiple() == 9600 and p.forecast(9600) == [9599 + 18365]
    tr = gr_emu.run_chain(y, p, PassThrough("sss"))
    assert len(tr.pss_calls) >= 50
    op = oracle.Pss(k, 4.0)
    buf = np.concatenate([np.zeros(960, np.complex64), y])
    pos, written, i_emit, n_lost = 960, 0, 0, 0
    for (r, nout, ncons, tags) in tr.pss_calls:
        assert r == pos - 960
        w_nout, w_ncons, w_out, w_rec = op.work(buf, pos)
        assert (nout, ncons) == (w_nout, w_ncons)
        lost = bool(w_rec["flags"] & oracle.F_TAG_LOST)
        n_lost += lost
        assert [(t.key, t.offset, t.value) for t in tags] == ([("tracking_lost", written, None)] if lost else [])
        if nout:
            assert np.array_equal(tr.pss_out[i_emit].view(np.uint32), w_out.view(np.uint32))
            i_emit += 1
            written += nout
        pos += ncons
    assert 15 <= n_lost <= 20 and i_emit > 40                    # track_after - 1 tagged half-frames before tracking starts (the 16th
    # window starts tracking before it emits); a decimated stream loses the first alignment to the filter's group delay and starts over
    assert p.tracking_score() == op.tracking_score() == 16.0 and p.max_psr() == op.max_psr() and p.mean_cfo() == op.mean_cfo()
    p.set_psr_threshold(1.0)                                     # the block itself does not clamp
    assert p.psr_threshold() == 1.0
    with pytest.raises(RuntimeError):
        lt.pss(3, 4.0)                                           # lib/pss_impl.cc:75-76
