"""The hier-block mirror `downlink_trigger_c` (python/downlink_trigger_c.py:13-73), its three host `mib` blocks and
`cellstore` on CPU, with the oracle's restated blocks standing in for the CUDA engine behind `Trigger`'s interface: the
reference's QA flowgraphs (python/qa_downlink_trigger_c.py:67-203) and the host logic around the engine -- message ports,
threshold clamp, tag offsets per chain, records, exit_on_success -- without a device.  tests/test_gpu_blocks.py runs the
same assertions on the real engine."""
import numpy as np
import pytest

from conftest import FIXTURES, load_fixture
import oracle_engine

NOF_PRB = {"6prb": 6, "25prb": 25, "50prb": 50, "100prb": 100}


@pytest.fixture()
def hier(oracle, monkeypatch):
    import ltetrigger_b200 as lt
    from ltetrigger_b200 import blocks
    monkeypatch.setattr(blocks, "Trigger", oracle_engine.make(oracle))
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


def test_cell_search_batch_cli(oracle, monkeypatch, tmp_path):
    """examples/cell_search_batch.py: several files as the streams of one engine, one JSON object per file -- a cell, a
    noise-only capture (NOT_FOUND), and the same cell again from an sc16 file through the restated conversion."""
    import json
    import os
    import sys
    import ltetrigger_b200 as lt
    from conftest import GOLDEN, ROOT
    from ltetrigger_b200 import synth
    eng = oracle_engine.make(oracle)
    monkeypatch.setattr(lt, "Trigger", eng)
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import cell_search_batch as cli
    frame = os.path.join(GOLDEN, "test_frames", "lte_frame_6prb_cellid_123")
    noise = str(tmp_path / "noise.fc32")
    synth.capture(0, 19200 * 12, snr_db=0.0, seed=9, noise_only=True).tofile(noise)
    out = cli.main(cli.parse(["-s", "1.92M", frame, noise, "--repeat", "--cut-off", "384000"]))
    a, b = (json.loads(o) for o in out)
    assert (a["status"], a["cell_id"], a["nof_prb"], a["cp_len"], os.path.basename(a["file"])) == \
        ("FOUND", 123, 6, "Normal", "lte_frame_6prb_cellid_123")
    assert b == {"status": "NOT_FOUND", "file": noise}
    assert eng.created[-1]["n_streams"] == 2 and eng.created[-1]["record_all"] is False and eng.created[-1]["keep_halfframes"] is True
    x = np.fromfile(frame, np.complex64)
    sc16 = str(tmp_path / "frame.sc16")
    synth.to_sc16(x[None, :])[0].tofile(sc16)
    out = cli.main(cli.parse(["-s", "1.92M", "--format", "sc16", sc16, "--repeat", "--cut-off", "384000"]))
    assert json.loads(out[0])["cell_id"] == 123 and eng.created[-1]["input_format"] == lt.FMT_SC16
