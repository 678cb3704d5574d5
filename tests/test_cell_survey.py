"""examples/cell_survey.py on CPU: the tool's host logic (segment plan, 50 ms passes, MIB confirmation, stitching, the
JSON it prints) with the oracle standing in for the CUDA engine behind `Trigger`'s interface.  The same scenario runs on
the real engine in tests/test_gpu_segments.py."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def two_cells_in_time():
    from ltetrigger_b200 import synth
    n = 19200 * 200
    a = synth.capture(301, n, seed=1, mib=dict(nof_prb=25, n_ports=1, phich_res=2))
    b = synth.capture(77, n, seed=2, ext_cp=True, mib=dict(nof_prb=50, n_ports=2, phich_res=1, h=(0.9 + 0.2j, -0.4 + 0.7j)))
    a[n // 2:] = 0
    b[:n // 2] = 0
    return (a + b + synth.capture(0, n, snr_db=10.0, seed=3, noise_only=True)).astype(np.complex64)


def test_cell_survey_with_the_oracle_as_engine(oracle, tmp_path, monkeypatch):
    import ltetrigger_b200 as lt
    from ltetrigger_b200 import _abi as A, shard

    x = two_cells_in_time()
    length = shard.plan_time_segments(len(x), 1, 8).length

    class OracleTrigger:
        """Collects the passes and answers the last one with the oracle's records of the whole segments; half-frames are
        raw slices (no carrier offset in this capture, so the CFO rotation the engine applies is negligible)."""

        made = []

        def __init__(self, n_streams, decim, psr_threshold, **kw):
            self.decim, self.thr, self.parts, self.last = decim, psr_threshold, [], np.zeros(0, A.WINDOW_REC)
            OracleTrigger.made.append(dict(kw, n_streams=n_streams, decim=decim))

        def process(self, chunk):
            self.parts.append(np.array(chunk))
            if sum(p.shape[1] for p in self.parts) == length:
                self.rows = np.concatenate(self.parts, 1)
                self.last = oracle.trigger_run(self.rows, decim=self.decim, psr_threshold=self.thr, conv_mode=oracle.CONV_OS)
            return self.last

        def fetch_halfframes(self, k):
            em = self.last[(self.last["flags"] & lt.F_EMIT) != 0]
            return np.stack([self.rows[int(r["stream"]), int(r["emit_start"]):int(r["emit_start"]) + 9600] for r in em])

        def close(self):
            pass

    monkeypatch.setattr(lt, "Trigger", OracleTrigger)
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import cell_survey
    path = str(tmp_path / "two_cells.fc32")
    x.tofile(path)
    cells = {c["cell_id"]: c for c in cell_survey.main(cell_survey.parse(["-s", "1.92M", "--segments", "8", path]))}
    assert sorted(cells) == [77, 301]
    c = cells[301]
    assert (c["nof_prb"], c["nof_tx_ports"], c["cp_len"], c["nof_phich_resources"], c["halfframes"]) == (25, 1, "Normal", "1", 167)
    assert abs(c["first_seen_s"] - 0.0885) < 1e-3 and abs(c["last_seen_s"] - 0.9985) < 1e-3
    c = cells[77]
    assert (c["nof_prb"], c["nof_tx_ports"], c["cp_len"], c["nof_phich_resources"], c["halfframes"]) == (50, 2, "Extended", "1/2", 177)
    assert abs(c["first_seen_s"] - 1.1066) < 1e-3 and abs(c["last_seen_s"] - 1.9866) < 1e-3
    # the sequential search of the whole capture tags the same half-frames of both cells
    seq = oracle.trigger_run(x[None, :], decim=1, psr_threshold=4.0, conv_mode=oracle.CONV_OS)
    for cell, n in ((301, 167), (77, 177)):
        assert int((((seq["flags"] & lt.F_CELL) != 0) & (seq["cell_id"] == cell)).sum()) == n

    # the same capture as a SigMF recording: rate and sample type come from the metadata, the options reach the engine
    from ltetrigger_b200 import sigmf
    sigmf.write(str(tmp_path / "two_cells"), x, 1.92e6, frequency=2.11e9)
    again = cell_survey.main(cell_survey.parse(["--segments", "8", "--frontend", "tc", "--full-scale", "8", str(tmp_path / "two_cells.sigmf-meta")]))
    assert sorted(c["cell_id"] for c in again) == [77, 301]
    kw = OracleTrigger.made[-1]
    assert (kw["n_streams"], kw["decim"], kw["input_format"], kw["frontend_mode"], kw["fc32_full_scale"]) == \
        (8, 1, lt.FMT_FC32, lt.FRONTEND_TC_INT, 8.0)
    assert OracleTrigger.made[0]["frontend_mode"] == lt.FRONTEND_FP32 and OracleTrigger.made[0]["fc32_full_scale"] == 0.0
    with pytest.raises(SystemExit):
        cell_survey.survey(cell_survey.parse(["--frontend", "tc", "-s", "1.92M", path]))          # fc32 + tc needs the range
    with pytest.raises(SystemExit):
        cell_survey.survey(cell_survey.parse([path]))                                             # raw file without a rate
