"""The drop-in block mirrors (blocks.py: pss, sss, downlink_trigger_c) driven like the reference's
QA flowgraphs (python/qa_downlink_trigger_c.py:67-203), call by call against the oracle's
restated blocks: produced/consumed counts, stream tags, emitted samples, accessors."""
import numpy as np
import pytest

from conftest import FIXTURES, load_fixture

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lt():
    import ltetrigger_b200 as lt
    if lt.device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU path")
    return lt


def search_rate(lt, name, seconds):
    x, decim, cell_id = load_fixture(name, seconds)
    if decim > 1:
        x = lt.kernel_decimate(x[None, :], decim)[0]        # rational_resampler_ccc(1, D) in front
    return x, cell_id


@pytest.mark.parametrize("name", ["6prb", "25prb"])
def test_pss_sss_blocks_call_by_call(lt, oracle, name):
    from ltetrigger_b200 import gr_emu
    x, cell_id = search_rate(lt, name, 0.4)
    k = cell_id % 3
    p, s = lt.pss(k, 4.0), lt.sss(k)
    assert p.history() == 9600 and p.output_multiple() == 9600 and s.output_multiple() == 9600
    tr = gr_emu.run_chain(x, p, s)
    assert len(tr.pss_calls) >= 70
    # oracle blocks driven over the same stream
    op, os_ = oracle.Pss(k, 4.0), oracle.Sss(k)
    buf = np.concatenate([np.zeros(960, np.complex64), x])
    pos, written, i_emit = 960, 0, 0
    for (r, nout, ncons, tags) in tr.pss_calls:
        assert r == pos - 960
        w_nout, w_ncons, w_out, w_rec = op.work(buf, pos)
        assert (nout, ncons) == (w_nout, w_ncons), (r, nout, ncons, w_nout, w_ncons)
        lost = bool(w_rec["flags"] & oracle.F_TAG_LOST)
        assert [(t.key, t.offset, t.value) for t in tags] == ([("tracking_lost", written, None)] if lost else [])
        if nout:
            assert np.array_equal(tr.pss_out[i_emit].view(np.uint32), w_out.view(np.uint32))
            s_out, s_rec = os_.work(w_out, lost)
            _, in_tags, out_tags = tr.sss_calls[i_emit]
            want = []
            if s_rec["flags"] & oracle.F_CELL:
                want = [("cell_id", written, int(s_rec["cell_id"])), ("cp_type", written, bool(s_rec["flags"] & oracle.F_CP_NORM))]
            assert [(t.key, t.offset, t.value) for t in out_tags] == want
            if lost or (s_rec["flags"] & oracle.F_CELL):
                assert np.array_equal(tr.sss_out[i_emit].view(np.uint32), w_out.view(np.uint32))   # pass-through
            else:
                assert tr.sss_out[i_emit] is None                                                   # :119-120
            i_emit += 1
            written += nout
        pos += ncons
    cells = {t.value for (_, _, out_tags) in tr.sss_calls for t in out_tags if t.key == "cell_id"}
    assert cells == {cell_id}
    # accessors (lib/pss_impl.h:95-100)
    assert p.tracking_score() == op.tracking_score() == 16.0
    assert p.max_psr() == op.max_psr() and p.mean_psr() == op.mean_psr() and p.mean_cfo() == op.mean_cfo()
    assert p.psr_threshold() == 4.0
    p.set_psr_threshold(1.0)                       # the block itself does not clamp
    assert p.psr_threshold() == 1.0


def test_sss_block_tdd(lt, oracle):
    """The standalone sss block with frame_type = TDD (not in the reference) on the half-frames the
    pss block aligns out of a TDD capture, call by call against the oracle's sss block."""
    from ltetrigger_b200 import synth
    x = synth.capture(205, 19200 * 16, snr_db=15.0, seed=3, tdd=True)       # tracking (and SSS) after 16 hits
    k = 205 % 3
    blk, ob, op = lt.sss(k, frame_type=lt.FRAME_TDD), oracle.Sss(k, frame_type=1), oracle.Pss(k, 3.0)
    buf = np.concatenate([np.zeros(960, np.complex64), x])
    pos, n, cells = 960, 0, []
    while pos - 960 + oracle.LOOKAHEAD <= len(x):
        nout, ncons, out, rec = op.work(buf, pos)
        if nout:
            lost = bool(rec["flags"] & oracle.F_TAG_LOST)
            _, want = ob.work(out, lost)
            blk._in_tags = [lt.tag_t(blk.nitems_read(0), "tracking_lost", None)] if lost else []
            assert blk.work(9600, [out], [np.zeros(9600, np.complex64)]) == 9600
            got = blk.last_record
            mask = lt.F_CELL | lt.F_CP_NORM | lt.F_SSS
            assert (int(got["flags"]) & mask) == (int(want["flags"]) & mask)
            for f in ("m0", "m1", "n_id_1", "cell_id"):
                assert got[f] == want[f], f
            assert np.float32(got["m0_val"]).tobytes() == np.float32(want["m0_val"]).tobytes()
            if int(got["flags"]) & lt.F_CELL:
                cells.append(int(got["cell_id"]))
            blk._nitems_read += 9600
            n += 1
        pos += ncons
    assert n >= 20 and len(cells) >= 4 and set(cells) == {205}


def test_pss_constructor_errors(lt):
    with pytest.raises(RuntimeError):
        lt.pss(3, 4.0)                             # lib/pss_impl.cc:75-76
    with pytest.raises(RuntimeError):
        lt.sss(7)


NOF_PRB = {"6prb": 6, "25prb": 25, "50prb": 50, "100prb": 100}


@pytest.mark.parametrize("name", list(FIXTURES))
def test_downlink_trigger_c_like_reference_qa(lt, name):
    """python/qa_downlink_trigger_c.py:67-203 as written: file_source(repeat) -> head(1 s) ->
    [rational_resampler_ccc(1, D)] -> downlink_trigger_c(psr_threshold=4, exit_on_success=True),
    message_debug on "track"; then the reference's six assertions on the first message."""
    x, cell_id = search_rate(lt, name, 1.0)
    trig = lt.downlink_trigger_c(psr_threshold=4, exit_on_success=True)
    assert trig.message_ports() == ["track", "drop"]
    tracked, dropped = [], []
    trig.msg_connect("track", tracked.append)
    trig.msg_connect("drop", dropped.append)
    tags = []
    for a in range(0, len(x), 96000):              # the scheduler hands over arbitrary chunks
        tags += trig.work(x[a:a + 96000])
    assert len(tracked) >= 1                       # self.assertTrue(self.msg_store.num_messages() >= 1)
    cell = tracked[0]
    assert cell["cell_id"] == cell_id              # _check_cell_id
    assert cell["cp_len"] == "Normal"              # _check_cp_len
    assert cell["nof_phich_resources"] == "1"      # _check_nof_phich_resources
    assert cell["nof_prb"] == NOF_PRB[name]        # _check_nof_prb
    assert cell["nof_tx_ports"] == 1               # _check_nof_tx_ports
    assert cell["phich_len"] == "Normal"           # _check_phich_len
    assert dropped == [] and len(tracked) == 1     # one cell, published once by its chain's mib
    assert [m.done for m in (trig.mib0, trig.mib1, trig.mib2)] == [k == cell_id % 3 for k in range(3)]
    ids = [(k, t.value) for k, t in tags if t.key == "cell_id"]
    assert {v for _, v in ids} == {cell_id} and {k for k, _ in ids} == {cell_id % 3}
    assert all(t.value is True for _, t in tags if t.key == "cp_type")
    assert trig.pss0.tracking_score() == (16.0 if cell_id % 3 == 0 else 0.0)
    # threshold clamp (python/downlink_trigger_c.py:63-73)
    trig.set_psr_threshold(0.5)
    assert trig.psr_threshold == 1.5 and trig.pss1.psr_threshold() == 1.5


def test_reference_snr_demo_screenshot_gpu(lt):
    """docs/gr_ltetrigger_snr_demo.png of the reference through the GPU hier block: -10.4 dB, threshold
    1.7 -> cell 123 tracked, 6 PRB, PHICH resources '1', normal CP (tests/conftest.py::snr_demo_capture)."""
    from conftest import snr_demo_capture
    y = snr_demo_capture(8.0, 0)
    trig = lt.downlink_trigger_c(psr_threshold=1.7, exit_on_success=True)
    tracked = []
    trig.msg_connect("track", tracked.append)
    for a in range(0, len(y) // 8 * 8, 96000):
        trig.work(y[a:a + 96000])
        if tracked:
            break
    assert len(tracked) >= 1
    cell = tracked[0]
    assert (cell["cell_id"], cell["nof_prb"], cell["nof_phich_resources"], cell["cp_len"]) == (123, 6, "1", "Normal")
    assert trig.pss0.tracking_score() > 0


def test_custom_mib_sink_and_drop(lt):
    """A custom mib stage sees every emitted half-frame with its tags; track / drop forwarding."""
    x, cell_id = search_rate(lt, "6prb", 0.3)
    x = np.concatenate([x, (np.random.default_rng(0).standard_normal((1200000, 2)) * 0.5).astype(np.float32).view(np.complex64)[:, 0]])
    trig = lt.downlink_trigger_c(psr_threshold=4)
    tracked, dropped = [], []
    trig.msg_connect("track", tracked.append)
    trig.msg_connect("drop", dropped.append)
    trig.work(x[:len(x) // 8 * 8])
    assert len(tracked) == 1 and tracked[0]["cell_id"] == cell_id and tracked[0]["nof_prb"] == 6
    assert dropped == tracked                      # signal gone: the identical object goes out on "drop"


@pytest.mark.parametrize("name,rate", [("6prb", "1.92M"), ("25prb", "7.68M"), ("50prb", "15.36M"), ("100prb", "30.72M")])
def test_cell_search_file_cli(lt, name, rate, capsys):
    """examples/test.sh:3-6: the CLI over the four fixtures with --repeat --time-out 1."""
    import importlib.util
    import json
    import os
    from conftest import GOLDEN, ROOT
    spec = importlib.util.spec_from_file_location("cell_search_file", os.path.join(ROOT, "examples", "cell_search_file.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    fname, _, cell_id = FIXTURES[name]
    args = cli.parse([os.path.join(GOLDEN, "test_frames", fname), "-s", rate, "--repeat", "--time-out", "5"])
    results = cli.main(args)
    out = capsys.readouterr().out
    assert out.startswith("Starting cell search... done.")
    assert len(results) == 1
    cell = json.loads(results[0])
    assert cell["status"] == "FOUND" and cell["cell_id"] == cell_id and cell["nof_prb"] == NOF_PRB[name]
    assert cell["cp_len"] == "Normal" and cell["nof_tx_ports"] == 1
    # --cut-off counts samples AFTER the resampler, as the reference's head block does
    # (examples/cell_search_file.py:47-48, 69-77): 200 ms of search-rate samples find the cell at every
    # rate; counted at the input rate they would be 200 ms / decim and find nothing at decim >= 4
    res = cli.main(cli.parse([os.path.join(GOLDEN, "test_frames", fname), "-s", rate, "--repeat", "--cut-off", "384000"]))
    assert json.loads(res[0])["status"] == "FOUND"
    res = cli.main(cli.parse([os.path.join(GOLDEN, "test_frames", fname), "-s", rate, "--repeat", "--cut-off", "96000"]))
    assert json.loads(res[0]) == {"status": "NOT_FOUND"}          # 50 ms: fewer than track_after windows
    # no cell: noise capture without --repeat runs to the end of the file
    noise = (np.random.default_rng(3).standard_normal((400000, 2)) * 0.3).astype(np.float32)
    path = os.path.join(str(os.environ.get("TMPDIR", "/tmp")), "ltb_noise_%d.fc32" % os.getpid())
    noise.tofile(path)
    try:
        res = cli.main(cli.parse([path, "-s", "1.92M"]))
        assert json.loads(res[0]) == {"status": "NOT_FOUND"}
        with pytest.raises(SystemExit):
            cli.main(cli.parse([path, "-s", "2M"]))          # not a multiple of 1.92 MHz
    finally:
        os.remove(path)


@pytest.mark.parametrize("name,rate", [("50prb", "15.36M"), ("100prb", "30.72M")])
def test_cell_search_file_cli_with_the_tensor_core_front_end(lt, name, rate):
    """The CLI with --frontend tc --full-scale 4: the fused resampler runs as exact-integer GEMMs on the tensor cores, the
    hier block with its host-side MIB decode finds the same cell with the same MIB as with the float32 resampler."""
    import importlib.util
    import json
    import os
    from conftest import GOLDEN, ROOT
    spec = importlib.util.spec_from_file_location("cell_search_file", os.path.join(ROOT, "examples", "cell_search_file.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    fname, _, cell_id = FIXTURES[name]
    path = os.path.join(GOLDEN, "test_frames", fname)
    tc = json.loads(cli.main(cli.parse([path, "-s", rate, "--repeat", "--time-out", "5", "--frontend", "tc", "--full-scale", "4"]))[0])
    fp = json.loads(cli.main(cli.parse([path, "-s", rate, "--repeat", "--time-out", "5"]))[0])
    assert tc["status"] == "FOUND" and tc["cell_id"] == cell_id and tc["nof_prb"] == NOF_PRB[name]
    tc.pop("tracking_start_time", None), fp.pop("tracking_start_time", None)      # wall clock of the first track message
    assert tc == fp
    with pytest.raises(SystemExit):
        cli.main(cli.parse([path, "-s", rate, "--repeat", "--time-out", "1", "--frontend", "tc"]))      # fc32 needs its range


def test_cell_search_batch_cli(lt, tmp_path):
    """examples/cell_search_batch.py: several captures as the streams of one engine; per file the
    reference's cell dictionary or NOT_FOUND (same fields as cell_search_file.py)."""
    import importlib.util
    import json
    import os
    import sys
    from conftest import GOLDEN, ROOT
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    spec = importlib.util.spec_from_file_location("cell_search_batch", os.path.join(ROOT, "examples", "cell_search_batch.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    fixture = os.path.join(GOLDEN, "test_frames", FIXTURES["25prb"][0])
    x = np.fromfile(fixture, np.complex64)
    rng = np.random.default_rng(9)
    noisy = np.tile(x, 8)
    noisy = (noisy + 0.3 * np.sqrt(np.mean(np.abs(x) ** 2)) * (rng.standard_normal(len(noisy)) + 1j * rng.standard_normal(len(noisy)))).astype(np.complex64)
    noise = (0.1 * (rng.standard_normal(len(noisy)) + 1j * rng.standard_normal(len(noisy)))).astype(np.complex64)
    f_noisy, f_noise = str(tmp_path / "noisy.fc32"), str(tmp_path / "noise.fc32")
    noisy.tofile(f_noisy)
    noise.tofile(f_noise)
    res = [json.loads(r) for r in cli.main(cli.parse([fixture, f_noisy, f_noise, "-s", "7.68M", "--repeat", "--cut-off", "7.68M"]))]
    assert [r["status"] for r in res] == ["FOUND", "FOUND", "NOT_FOUND"]
    for r in res[:2]:
        assert (r["cell_id"], r["nof_prb"], r["cp_len"], r["nof_tx_ports"], r["nof_phich_resources"]) == (124, 25, "Normal", 1, "1")
    assert [os.path.basename(r["file"]) for r in res] == [os.path.basename(fixture), "noisy.fc32", "noise.fc32"]
    # the same captures as sc16 files
    f16 = str(tmp_path / "noisy.sc16")
    s = 32767.0 / (4 * np.abs(noisy).max())
    np.stack([np.round(noisy.real * s), np.round(noisy.imag * s)], axis=-1).astype(np.int16).tofile(f16)
    res = [json.loads(r) for r in cli.main(cli.parse([f16, "-s", "7.68M", "--format", "sc16", "--repeat", "--cut-off", "7.68M"]))]
    assert res[0]["status"] == "FOUND" and res[0]["cell_id"] == 124 and res[0]["nof_prb"] == 25


def test_cell_search_file_cli_on_synthetic_75prb_two_port_cell(lt, tmp_path):
    """The CLI at a rate the bundled frames do not cover: a synthetic 15 MHz cell (75 PRB, 23.04 Msps,
    decimate by 12) with two antenna ports -> FOUND with the MIB the transmitter encoded."""
    import importlib.util
    import json
    import os
    from conftest import ROOT
    from ltetrigger_b200 import synth
    spec = importlib.util.spec_from_file_location("cell_search_file", os.path.join(ROOT, "examples", "cell_search_file.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    x = synth.capture(332, 19200 * 12 * 14, snr_db=10.0, decim=12, seed=8,
                      mib=dict(nof_prb=75, n_ports=2, phich_res=3, sfn0=8, h=(0.7 - 0.2j, 0.3 + 0.8j)))
    path = str(tmp_path / "cell332_75prb.fc32")
    x.tofile(path)
    results = cli.main(cli.parse([path, "-s", "23.04M"]))
    cell = json.loads(results[0])
    assert cell["status"] == "FOUND"
    assert (cell["cell_id"], cell["nof_prb"], cell["nof_tx_ports"], cell["nof_phich_resources"]) == (332, 75, 2, "2")
