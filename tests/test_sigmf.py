"""SigMF ingest (scope row f4): metadata parsing on the CPU, and the reference's CLI reading SigMF recordings of the
bundled frames -- as fc32 and in the integer wire formats, which go to the GPU without a host conversion."""
import importlib.util
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def _cli():
    spec = importlib.util.spec_from_file_location("cell_search_file", os.path.join(ROOT, "examples", "cell_search_file.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    return cli


def test_sigmf_round_trip_and_errors(tmp_path):
    from ltetrigger_b200 import sigmf, FMT_FC32, FMT_SC16, FMT_SC8
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(1000) + 1j * rng.standard_normal(1000)).astype(np.complex64)
    for samples, fmt, dt in ((x, FMT_FC32, "cf32_le"),
                             (rng.integers(-30000, 30000, (1000, 2)).astype(np.int16), FMT_SC16, "ci16_le"),
                             (rng.integers(-128, 128, (1000, 2)).astype(np.int8), FMT_SC8, "ci8")):
        meta, data = sigmf.write(str(tmp_path / ("rec_" + dt)), samples, 15.36e6, frequency=751e6)
        assert meta.endswith(".sigmf-meta") and data.endswith(".sigmf-data")
        for name in (meta, data, meta[:-len(".sigmf-meta")]):
            assert sigmf.is_sigmf(name)
            rec = sigmf.load(name)
            assert rec["input_format"] == fmt and rec["datatype"] == dt
            assert rec["sample_rate"] == 15.36e6 and rec["frequency"] == 751e6
            assert np.array_equal(np.asarray(rec["samples"]), samples)
    assert not sigmf.is_sigmf(str(tmp_path / "plain.fc32"))
    # what the search cannot take is refused with a message, not mis-read
    base = str(tmp_path / "bad")
    sigmf.write(base, x, 1.92e6)
    meta = json.load(open(base + ".sigmf-meta"))
    for key, val in (("core:datatype", "rf32_le"), ("core:datatype", "cf32_be"), ("core:datatype", "cu8"), ("core:num_channels", 2)):
        m = json.loads(json.dumps(meta))
        m["global"][key] = val
        json.dump(m, open(base + ".sigmf-meta", "w"))
        with pytest.raises(sigmf.SigMFError):
            sigmf.load(base)
    m = json.loads(json.dumps(meta))
    del m["global"]["core:sample_rate"]
    json.dump(m, open(base + ".sigmf-meta", "w"))
    with pytest.raises(sigmf.SigMFError):
        sigmf.load(base)
    # a header in front of the samples (core:header_bytes of the first capture) is skipped
    m = json.loads(json.dumps(meta))
    m["captures"][0]["core:header_bytes"] = 16
    json.dump(m, open(base + ".sigmf-meta", "w"))
    assert np.array_equal(np.asarray(sigmf.load(base)["samples"]), x[2:])


@pytest.mark.gpu
@pytest.mark.parametrize("name,fname,rate,cell_id,prb,kind", [
    ("50prb", "lte_frame_50prb_cellid_125", 15.36e6, 125, 50, "ci16_le"),
    ("100prb", "lte_frame_100prb_cellid_369", 30.72e6, 369, 100, "cf32_le"),
    ("25prb", "lte_frame_25prb_cellid_124", 7.68e6, 124, 25, "ci8"),
])
def test_cell_search_file_cli_reads_sigmf(tmp_path, name, fname, rate, cell_id, prb, kind):
    import ltetrigger_b200 as lt
    from ltetrigger_b200 import sigmf, synth
    if lt.device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU path")
    x = np.fromfile(os.path.join(GOLDEN, "test_frames", fname), np.complex64)
    samples = x if kind == "cf32_le" else synth.to_sc16(x[None])[0] if kind == "ci16_le" else synth.to_sc8(x[None])[0]
    base = str(tmp_path / name)
    sigmf.write(base, samples, rate, frequency=739e6)
    cli = _cli()
    res = cli.main(cli.parse([base + ".sigmf-meta", "--repeat", "--time-out", "5"]))          # no -s: the metadata has it
    cell = json.loads(res[0])
    assert cell["status"] == "FOUND" and cell["cell_id"] == cell_id and cell["nof_prb"] == prb and cell["cp_len"] == "Normal"
    with pytest.raises(SystemExit):
        cli.main(cli.parse([base + ".sigmf-data", "-s", "1.92M", "--repeat", "--time-out", "1"]))   # contradicts the metadata
