"""Time-segment sharding of one capture on the CUDA engine (shard.plan_time_segments / cut_segments /
stitch_segments): the segments run as the streams of one engine; every segment's records are bit-identical to the
oracle's search of that segment, and the stitched list holds exactly the cell-tagged half-frames of the sequential
search of the whole capture."""
import numpy as np
import pytest

from conftest import assert_recs_equal, load_fixture

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lt():
    import ltetrigger_b200 as lt
    if lt.device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU path")
    return lt


def _sorted(r):
    return r[np.lexsort((r["win_index"], r["n_id_2"], r["stream"]))]


def test_synthetic_capture_in_six_segments(lt, oracle):
    from ltetrigger_b200 import shard, synth
    decim = 4
    x = synth.capture(301, 19200 * decim * 100, snr_db=6.0, decim=decim, seed=5, cfo_hz=800.0)
    plan = shard.plan_time_segments(len(x), decim, 6)
    rows = shard.cut_segments(x, plan)
    trig = lt.Trigger(n_streams=plan.n_segments, decim=decim, psr_threshold=4.0, max_chunk=96000 * decim, corr_mode=lt.CORR_FFT)
    got = _sorted(trig.run(rows, chunk=96000 * decim))
    trig.close()
    want = oracle.trigger_run(rows, decim=decim, psr_threshold=4.0, conv_mode=oracle.CONV_OS)
    assert_recs_equal(got, want)
    st = shard.stitch_segments(got, plan)
    trig = lt.Trigger(n_streams=1, decim=decim, psr_threshold=4.0, max_chunk=96000 * decim, corr_mode=lt.CORR_FFT)
    seq = trig.run(x[None, :], chunk=96000 * decim)
    trig.close()
    c_seq, c_st = (r[(r["flags"] & lt.F_CELL) != 0] for r in (_sorted(seq), st))
    assert len(c_seq) > 150 and c_st["emit_start"].tolist() == c_seq["emit_start"].tolist()
    assert set(c_st["cell_id"].tolist()) == {301}


def test_100prb_fixture_sc16_tensor_core_front_end_in_segments(lt, oracle):
    """1 s of the 100 PRB frame as sc16 at 30.72 Msps, seven segments of 0.26 s through the integer tensor-core front end."""
    from ltetrigger_b200 import shard, synth
    x, decim, cell = load_fixture("100prb", seconds=1.0)
    iq = synth.to_sc16(x[None, :])[0]
    plan = shard.plan_time_segments(len(iq), decim, 16)
    assert plan.n_segments == 7
    rows = shard.cut_segments(iq, plan)
    trig = lt.Trigger(n_streams=plan.n_segments, decim=decim, psr_threshold=4.0, max_chunk=96000 * decim,
                      input_format=lt.FMT_SC16, corr_mode=lt.CORR_FFT, frontend_mode=lt.FRONTEND_TC_INT)
    got = _sorted(trig.run(rows, chunk=96000 * decim))
    trig.close()
    want = oracle.trigger_run(rows, decim=decim, fmt=lt.FMT_SC16, psr_threshold=4.0, conv_mode=oracle.CONV_OS | oracle.FRONT_TCINT)
    assert_recs_equal(got, want)
    st = shard.stitch_segments(got, plan)
    cells = st[(st["flags"] & lt.F_CELL) != 0]
    assert set(cells["cell_id"].tolist()) == {cell} and len(cells) > 150
    assert set(np.diff(cells["emit_start"]).tolist()) == {9600}          # no tagged half-frame lost or doubled at a boundary


def test_cell_survey_cli_two_cells_in_time(lt, tmp_path):
    """examples/cell_survey.py on a 2 s capture: cell 301 (25 PRB, one port) on air during the first second, cell 77
    (50 PRB, two ports, extended CP) during the second; eight segments side by side.  Both are listed with their MIB and
    with the times the sequential search reports (first tagged half-frame 88.5 ms after a cell appears: track_after
    windows and one more)."""
    import os
    import sys
    from conftest import ROOT
    from test_cell_survey import two_cells_in_time
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import cell_survey
    x = two_cells_in_time()
    path = str(tmp_path / "two_cells.fc32")
    x.tofile(path)
    cells = cell_survey.main(cell_survey.parse(["-s", "1.92M", "--segments", "8", path]))
    by_id = {c["cell_id"]: c for c in cells}
    assert sorted(by_id) == [77, 301], sorted(by_id)                    # sss tags without a MIB behind them are not cells
    c = by_id[301]
    assert (c["nof_prb"], c["nof_tx_ports"], c["cp_len"], c["nof_phich_resources"]) == (25, 1, "Normal", "1")
    assert c["halfframes"] == 167 and abs(c["first_seen_s"] - 0.0885) < 1e-3 and abs(c["last_seen_s"] - 0.9985) < 1e-3
    c = by_id[77]
    assert (c["nof_prb"], c["nof_tx_ports"], c["cp_len"], c["nof_phich_resources"]) == (50, 2, "Extended", "1/2")
    assert c["halfframes"] == 177 and abs(c["first_seen_s"] - 1.1066) < 1e-3 and abs(c["last_seen_s"] - 1.9866) < 1e-3
