"""LTB_FRONTEND_TC_INT: the exact-integer tensor-core front end (tcgen05.mma kind::i8, sc16 input at
decim 16) against the oracle's int64 restatement (ORC_FRONT_TCINT): bit for bit, in ragged chunks, at
the digit extremes, through the engine, and within the north_star's tolerance of the canonical float32
front end.  Last in the GPU suite on purpose: a protocol error in this kernel traps (its watchdogs never
hang) and would take the CUDA context of the test process with it."""
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


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("fmt", ["sc16", "sc8"])
@pytest.mark.parametrize("n_out,chunk", [(976 * 3, None), (5000, None), (20000, 128 * 37), (20000, 128 * 2), (61 * 16 * 5 + 8, 128 * 125)])
def test_tc_decimator_bit_exact(lt, oracle, n_out, chunk, fmt):
    """Random full-scale integer streams; chunk sizes that are not multiples of a 256-sample row, smaller
    than the 768-sample history, and tile-aligned; every output equal to the int64 evaluation."""
    rng = np.random.default_rng(n_out)
    n = n_out * 16
    lo, hi, dt = (-32768, 32767, np.int16) if fmt == "sc16" else (-128, 127, np.int8)
    x = rng.integers(lo, hi + 1, size=(3, n, 2)).astype(dt)
    x[1, :, :] = hi                         # digit extremes: all-max and all-min streams
    x[2, :, 0] = lo
    got = lt.kernel_decimate_tc(x, chunk=chunk)
    ref = oracle.decimate_tcint_sc16 if fmt == "sc16" else oracle.decimate_tcint_sc8
    for s in range(3):
        want = ref(x[s])
        assert np.array_equal(_bits(got[s]), _bits(want)), (s, int(np.argmax(_bits(got[s]) != _bits(want))))


def test_tc_decimator_within_tolerance_of_float32_front_end(lt, oracle):
    """north_star: magnitudes within 1e-4 relative.  The integer front end differs from the canonical
    float32 decimator only by tap quantisation (2^-28) and one rounding: ~3e-7 of full scale."""
    rng = np.random.default_rng(7)
    x = rng.integers(-20000, 20000, size=(2, 16 * 4000, 2)).astype(np.int16)
    got = lt.kernel_decimate_tc(x)
    ref = lt.kernel_decimate(x, 16, lt.FMT_SC16)
    assert np.abs(got - ref).max() < 1e-5 * np.abs(ref).max()


@pytest.mark.parametrize("fmt", ["sc16", "sc8"])
@pytest.mark.parametrize("corr", ["fft", "direct"])
def test_tc_front_end_through_the_engine(lt, oracle, corr, fmt):
    """The 100 PRB fixture and synthetic cells as sc16 at 30.72 Msps through the whole chain with the
    integer front end, in ragged chunks: records bit-identical to the oracle in the same mode, the same
    decisions as the float32 front end, the reference's known cell id."""
    from ltetrigger_b200 import synth
    x, decim, cell_id = load_fixture("100prb", 0.25)
    rows = [x, synth.capture(77, len(x), snr_db=3.0, decim=16, seed=5, cfo_hz=1500.0),
            synth.capture(300, len(x), snr_db=0.0, decim=16, seed=6, noise_only=True)]
    iq = synth.to_sc16(np.stack(rows)) if fmt == "sc16" else synth.to_sc8(np.stack(rows))
    code = lt.FMT_SC16 if fmt == "sc16" else lt.FMT_SC8
    mode = lt.CORR_FFT if corr == "fft" else lt.CORR_DIRECT
    conv = (oracle.CONV_OS if corr == "fft" else oracle.CONV_DIRECT) | oracle.FRONT_TCINT
    chunk = 16 * 8 * 4001
    trig = lt.Trigger(n_streams=3, decim=16, max_chunk=chunk, input_format=code, corr_mode=mode,
                      frontend_mode=lt.FRONTEND_TC_INT)
    got = trig.run(iq, chunk=chunk)
    trig.close()
    want = oracle.trigger_run(iq, decim=16, fmt=code, conv_mode=conv)
    assert_recs_equal(got, want)
    cells = got[(got["flags"] & lt.F_CELL) != 0]
    assert set(cells[cells["stream"] == 0]["cell_id"].tolist()) == {cell_id}
    assert set(cells[cells["stream"] == 1]["cell_id"].tolist()) == {77}
    # the float32 front end on the same input: same windows, flags, peaks and cell ids; magnitudes within 1e-4
    ref = lt.Trigger(n_streams=3, decim=16, max_chunk=chunk, input_format=code, corr_mode=mode)
    fp = ref.run(iq, chunk=chunk)
    ref.close()
    sel = (got["stream"] < 2)                                  # noise-only stream: argmax of noise may move
    for f in ("win_start", "emit_start", "flags", "peak_pos", "m0", "m1", "n_id_1", "cell_id"):
        assert (got[f][sel] == fp[f][sel]).all(), f
    np.testing.assert_allclose(got["psr"][sel], fp["psr"][sel], rtol=1e-4)
    np.testing.assert_allclose(got["peak_value"][sel], fp["peak_value"][sel], rtol=1e-4)


@pytest.mark.parametrize("n_out,chunk", [(976 * 3, None), (5000, None), (20000, 128 * 37), (20000, 128 * 2), (61 * 16 * 5 + 8, 128 * 125)])
def test_tc_decimator_fc32_fixed_point_bit_exact(lt, oracle, n_out, chunk):
    """fc32 input taken as 23-bit fixed point over +-full_scale: gaussian samples, streams pinned at the two
    saturation levels and beyond them, NaN / Inf / denormal samples; every output equal to the oracle's
    int64 evaluation of the same quantised samples, in ragged chunks."""
    rng = np.random.default_rng(n_out + 1)
    n = n_out * 16
    fs = 3.0
    x = (rng.standard_normal((4, n)) + 1j * rng.standard_normal((4, n))).astype(np.complex64) * np.float32(0.6)
    x[1] = fs * (1 + 1j)                                   # upper saturation level exactly
    x[2] = -5 * fs + 0j                                    # far below the range: saturates
    odd = np.array([np.nan, np.inf, -np.inf, 1e-42, -0.0, fs, -fs, fs * (1 - 2.0 ** -23)], np.float32)
    x[3, :4096].real = np.tile(odd, 512)
    x[3, 4096:8192].imag = np.tile(odd[::-1], 512)
    got = lt.kernel_decimate_tc(x, chunk=chunk, full_scale=fs)
    for s in range(4):
        want = oracle.decimate_tcint_fc32(x[s], fs)
        assert np.array_equal(_bits(got[s]), _bits(want)), (s, int(np.argmax(_bits(got[s]) != _bits(want))))


def test_tc_decimator_fc32_within_tolerance_of_float32_front_end(lt, oracle):
    """north_star: magnitudes within 1e-4 relative.  With the signal's rms at 1/8 of the declared range the
    fixed-point grid (2^-23 of the range per sample) leaves ~1e-6 of the output's rms."""
    rng = np.random.default_rng(8)
    x = (rng.standard_normal((2, 16 * 4000)) + 1j * rng.standard_normal((2, 16 * 4000))).astype(np.complex64)
    got = lt.kernel_decimate_tc(x, full_scale=8.0)
    ref = lt.kernel_decimate(x, 16, lt.FMT_FC32)
    assert np.abs(got - ref).max() < 2e-5 * np.sqrt(np.mean(np.abs(ref) ** 2))


@pytest.mark.parametrize("corr", ["fft", "direct"])
def test_tc_front_end_fc32_through_the_engine(lt, oracle, corr):
    """The 100 PRB fixture and synthetic cells as fc32 at 30.72 Msps through the whole chain with the fixed-point
    tensor-core front end, in ragged chunks: records bit-identical to the oracle in the same mode, the same
    decisions as the float32 front end and magnitudes within 1e-4 of it."""
    from ltetrigger_b200 import synth
    x, decim, cell_id = load_fixture("100prb", 0.25)
    rows = [x, synth.capture(77, len(x), snr_db=3.0, decim=16, seed=5, cfo_hz=1500.0),
            synth.capture(300, len(x), snr_db=0.0, decim=16, seed=6, noise_only=True)]
    iq = np.stack(rows).astype(np.complex64)
    fs = float(8 * np.sqrt(np.mean(np.abs(iq) ** 2)))
    mode = lt.CORR_FFT if corr == "fft" else lt.CORR_DIRECT
    conv = (oracle.CONV_OS if corr == "fft" else oracle.CONV_DIRECT) | oracle.FRONT_TCINT
    chunk = 16 * 8 * 4001
    trig = lt.Trigger(n_streams=3, decim=16, max_chunk=chunk, input_format=lt.FMT_FC32, corr_mode=mode,
                      frontend_mode=lt.FRONTEND_TC_INT, fc32_full_scale=fs)
    got = trig.run(iq, chunk=chunk)
    trig.close()
    want = oracle.trigger_run(iq, decim=16, fmt=lt.FMT_FC32, conv_mode=conv, fc32_full_scale=fs)
    assert_recs_equal(got, want)
    cells = got[(got["flags"] & lt.F_CELL) != 0]
    assert set(cells[cells["stream"] == 0]["cell_id"].tolist()) == {cell_id}
    assert set(cells[cells["stream"] == 1]["cell_id"].tolist()) == {77}
    ref = lt.Trigger(n_streams=3, decim=16, max_chunk=chunk, input_format=lt.FMT_FC32, corr_mode=mode)
    fp = ref.run(iq, chunk=chunk)
    ref.close()
    sel = (got["stream"] < 2)
    for f in ("win_start", "emit_start", "flags", "peak_pos", "m0", "m1", "n_id_1", "cell_id"):
        assert (got[f][sel] == fp[f][sel]).all(), f
    np.testing.assert_allclose(got["psr"][sel], fp["psr"][sel], rtol=1e-4)
    np.testing.assert_allclose(got["peak_value"][sel], fp["peak_value"][sel], rtol=1e-4)


@pytest.mark.parametrize("fmt,decim", [("fc32", 2), ("fc32", 4), ("fc32", 8), ("fc32", 12), ("sc16", 4), ("sc16", 8), ("sc16", 12), ("sc8", 8),
                                       ("fc32", 24), ("sc16", 24), ("sc8", 24), ("fc32", 32), ("sc16", 32), ("sc8", 32)])
@pytest.mark.parametrize("n_out,chunk_outs", [(976 * 2 + 40, None), (9000, 8 * 37), (9000, 8 * 1), (61 * 16 * 3 + 8, 8 * 250)])
def test_tc_decimator_other_rates_bit_exact(lt, oracle, fmt, decim, n_out, chunk_outs):
    """The LTE sampling rates below 30.72 Msps (decimation 2, 4, 8, 12): the same kernel with 16 D samples per row, its
    k-steps meeting one, two or three tap tables (the phase of a k-step's first sample within an output); random and
    extreme streams in ragged chunks, every output equal to the oracle's int64 evaluation."""
    rng = np.random.default_rng(n_out + decim)
    n = n_out * decim
    chunk = None if chunk_outs is None else chunk_outs * decim
    if fmt == "fc32":
        fs = 3.0
        x = (rng.standard_normal((3, n)) + 1j * rng.standard_normal((3, n))).astype(np.complex64) * np.float32(0.6)
        x[1] = fs * (1 - 1j)
        x[2, ::7] = 10 * fs
        got = lt.kernel_decimate_tc(x, chunk=chunk, full_scale=fs, decim=decim)
        ref = lambda v: oracle.decimate_tcint_fc32(v, fs, decim)
    else:
        lo, hi, dt = (-32768, 32767, np.int16) if fmt == "sc16" else (-128, 127, np.int8)
        x = rng.integers(lo, hi + 1, size=(3, n, 2)).astype(dt)
        x[1, :, :] = hi
        x[2, :, 0] = lo
        got = lt.kernel_decimate_tc(x, chunk=chunk, decim=decim)
        ref = (lambda v: oracle.decimate_tcint_sc16(v, decim)) if fmt == "sc16" else (lambda v: oracle.decimate_tcint_sc8(v, decim))
    for s in range(3):
        want = ref(x[s])
        assert np.array_equal(_bits(got[s]), _bits(want)), (s, int(np.argmax(_bits(got[s]) != _bits(want))))


@pytest.mark.parametrize("fmt,decim", [("sc16", 16), ("fc32", 16), ("sc8", 8), ("fc32", 2), ("sc16", 12)])
def test_tc_decimator_smallest_chunks_and_heavy_clipping(lt, oracle, fmt, decim):
    """Calls of 8 D samples -- half a row: the tensor map covers no whole row and every row of the tile comes from the
    carried history or sample by sample -- and, for fc32, a declared range of a third of the signal's rms (most samples
    clip): still every output equal to the oracle's."""
    rng = np.random.default_rng(decim)
    n = 8 * decim * 41
    if fmt == "fc32":
        x = (rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n))).astype(np.complex64)
        got = lt.kernel_decimate_tc(x, chunk=8 * decim, full_scale=0.33, decim=decim)
        want = [oracle.decimate_tcint_fc32(x[s], 0.33, decim) for s in range(2)]
    else:
        lo, hi, dt = (-32768, 32767, np.int16) if fmt == "sc16" else (-128, 127, np.int8)
        x = rng.integers(lo, hi + 1, size=(2, n, 2)).astype(dt)
        got = lt.kernel_decimate_tc(x, chunk=8 * decim, decim=decim)
        want = [(oracle.decimate_tcint_sc16 if fmt == "sc16" else oracle.decimate_tcint_sc8)(x[s], decim) for s in range(2)]
    for s in range(2):
        assert np.array_equal(_bits(got[s]), _bits(want[s])), (s, int(np.argmax(_bits(got[s]) != _bits(want[s]))))


@pytest.mark.parametrize("name,fmt", [("50prb", "sc16"), ("25prb", "fc32"), ("50prb", "sc8"), ("25prb", "sc16")])
def test_tc_front_end_other_rates_through_the_engine(lt, oracle, name, fmt):
    """The 50 PRB (15.36 Msps, D = 8) and 25 PRB (7.68 Msps, D = 4) fixtures through the whole chain with the integer front
    end: records bit-identical to the oracle in the same mode, the reference's cell id, the same decisions as the float32
    front end."""
    from ltetrigger_b200 import synth
    x, decim, cell_id = load_fixture(name, 0.25)
    n = len(x) // (8 * decim) * (8 * decim)
    x = x[:n]
    code = {"fc32": lt.FMT_FC32, "sc16": lt.FMT_SC16, "sc8": lt.FMT_SC8}[fmt]
    iq = x[None] if fmt == "fc32" else synth.to_sc16(x[None]) if fmt == "sc16" else synth.to_sc8(x[None])
    fs = float(8 * np.sqrt(np.mean(np.abs(x) ** 2))) if fmt == "fc32" else 0.0
    chunk = decim * 8 * 3001
    trig = lt.Trigger(n_streams=1, decim=decim, max_chunk=chunk, input_format=code, corr_mode=lt.CORR_FFT,
                      frontend_mode=lt.FRONTEND_TC_INT, fc32_full_scale=fs)
    got = trig.run(iq, chunk=chunk)
    trig.close()
    want = oracle.trigger_run(iq, decim=decim, fmt=code, conv_mode=oracle.CONV_OS | oracle.FRONT_TCINT, fc32_full_scale=fs)
    assert_recs_equal(got, want)
    cells = got[(got["flags"] & lt.F_CELL) != 0]
    assert set(cells["cell_id"].tolist()) == {cell_id}
    ref = lt.Trigger(n_streams=1, decim=decim, max_chunk=chunk, input_format=code, corr_mode=lt.CORR_FFT)
    fp = ref.run(iq, chunk=chunk)
    ref.close()
    for f in ("win_start", "emit_start", "flags", "m0", "m1", "n_id_1", "cell_id"):
        assert (got[f] == fp[f]).all(), f
    over = (got["flags"] & lt.F_OVER) != 0
    assert (got["peak_pos"][over] == fp["peak_pos"][over]).all()
    np.testing.assert_allclose(got["psr"][over], fp["psr"][over], rtol=1e-4)


@pytest.mark.parametrize("fmt,decim", [("sc16", 32), ("fc32", 32), ("sc8", 24), ("fc32", 12)])
def test_tc_front_end_wide_rates_through_the_engine(lt, oracle, fmt, decim):
    """Synthetic cells at 61.44 / 46.08 / 23.04 Msps (no bundled frame has these rates) through the whole chain with the
    integer front end, two streams in ragged chunks: records bit-identical to the oracle in the same mode, the right cell
    ids, the same decisions as the float32 front end (which at these rates is the general kernel)."""
    from ltetrigger_b200 import synth
    n = 19200 * decim * 12
    rows = [synth.capture(211, n, snr_db=6.0, decim=decim, seed=3, cfo_hz=700.0),
            synth.capture(88, n, snr_db=2.0, decim=decim, seed=4)]
    x = np.stack(rows)
    code = {"fc32": lt.FMT_FC32, "sc16": lt.FMT_SC16, "sc8": lt.FMT_SC8}[fmt]
    iq = x if fmt == "fc32" else synth.to_sc16(x) if fmt == "sc16" else synth.to_sc8(x)
    fs = 6.0 if fmt == "fc32" else 0.0
    chunk = decim * 8 * 2777
    trig = lt.Trigger(n_streams=2, decim=decim, max_chunk=chunk, input_format=code, corr_mode=lt.CORR_FFT,
                      frontend_mode=lt.FRONTEND_TC_INT, fc32_full_scale=fs)
    got = trig.run(iq, chunk=chunk)
    trig.close()
    want = oracle.trigger_run(iq, decim=decim, fmt=code, conv_mode=oracle.CONV_OS | oracle.FRONT_TCINT, fc32_full_scale=fs)
    assert_recs_equal(got, want)
    cells = got[(got["flags"] & lt.F_CELL) != 0]
    assert set(cells[cells["stream"] == 0]["cell_id"].tolist()) == {211}
    assert set(cells[cells["stream"] == 1]["cell_id"].tolist()) == {88}
    ref = lt.Trigger(n_streams=2, decim=decim, max_chunk=chunk, input_format=code, corr_mode=lt.CORR_FFT)
    fp = ref.run(iq, chunk=chunk)
    ref.close()
    for f in ("win_start", "emit_start", "flags", "m0", "m1", "n_id_1", "cell_id"):
        assert (got[f] == fp[f]).all(), f
    over = (got["flags"] & lt.F_OVER) != 0
    assert (got["peak_pos"][over] == fp["peak_pos"][over]).all()
    np.testing.assert_allclose(got["psr"][over], fp["psr"][over], rtol=1e-4)


def test_tc_front_end_argument_checks(lt):
    with pytest.raises(lt.LtbError):
        lt.Trigger(n_streams=1, decim=16, input_format=lt.FMT_FC32, frontend_mode=lt.FRONTEND_TC_INT)   # fc32 needs its range
    with pytest.raises(lt.LtbError):
        lt.Trigger(n_streams=1, decim=16, input_format=lt.FMT_FC32, frontend_mode=lt.FRONTEND_TC_INT, fc32_full_scale=-1.0)
    with pytest.raises(lt.LtbError):
        lt.Trigger(n_streams=1, decim=2, input_format=lt.FMT_SC16, frontend_mode=lt.FRONTEND_TC_INT)    # a 16-output row would be 128 bytes
    with pytest.raises(lt.LtbError):
        lt.Trigger(n_streams=1, decim=5, input_format=lt.FMT_FC32, frontend_mode=lt.FRONTEND_TC_INT, fc32_full_scale=1.0)
