"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on
the same inputs.  Integer/index/flag fields and every float are compared bit-exactly (the
kernels evaluate the oracle's canonical expression trees); the north_star's 1e-4 relative
tolerance on magnitudes applies to the oracle's reference-class FFT mode, checked too."""
import numpy as np
import pytest

from conftest import FIXTURES, assert_recs_equal, load_fixture

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lt():
    import ltetrigger_b200 as lt
    if lt.device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU path")
    return lt


def rand_c64(rng, *shape):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)


# ---- kernel level ---------------------------------------------------------------------
@pytest.mark.parametrize("n", [8, 2048, 2056, 30000])
def test_pss_corr_kernel_bit_exact(lt, oracle, n):
    rng = np.random.default_rng(n)
    x = rand_c64(rng, 3, n)
    got = lt.kernel_pss_corr(x)
    for s in range(3):
        for r in range(3):
            want = oracle.pss_corr_stream(x[s], r)
            assert np.array_equal(got[s, r].view(np.uint32), want.view(np.uint32)), (s, r)


def test_pss_corr_kernel_linearity_full_size(lt):
    """BASELINE-size property check (no oracle): correlation amplitude scales linearly, so
    power scales by exactly 4 when the input is doubled (power-of-two scaling is exact)."""
    rng = np.random.default_rng(5)
    x = rand_c64(rng, 4, 1 << 20)
    p1 = lt.kernel_pss_corr(x)
    p2 = lt.kernel_pss_corr(2 * x)
    assert np.array_equal(p2, 4 * p1)


@pytest.mark.parametrize("n", [896, 3000, 30000])
def test_pss_corr_fft_kernel_bit_exact(lt, oracle, n):
    """LTB_CORR_FFT: overlap-save blocks against the oracle's ORC_CONV_OS restatement, bit for bit,
    and against the direct form within the north_star's tolerance."""
    rng = np.random.default_rng(n)
    x = rand_c64(rng, 3, n)
    got = lt.kernel_pss_corr_fft(x)
    for s in range(3):
        want = oracle.pss_corr_os(x[s])
        assert got[s].shape == want.shape
        assert np.array_equal(got[s].view(np.uint32), want.view(np.uint32)), s
        for r in range(3):
            d = oracle.pss_corr_stream(x[s], r)[:want.shape[1]]
            assert np.abs(got[s, r] - d).max() < 1e-4 * d.max()


def _decim_case(oracle, rng, decim, fmt, n):
    if fmt == 0:
        x = rand_c64(rng, 2, n)
        want = [oracle.decimate(x[s], decim) for s in range(2)]
    elif fmt == 1:
        x = rng.integers(-32768, 32767, size=(2, n, 2), dtype=np.int16)
        want = [oracle.decimate(oracle.sc16_to_fc32(x[s]), decim) for s in range(2)]
    else:
        x = rng.integers(-128, 127, size=(2, n, 2), dtype=np.int8)
        want = [oracle.decimate(oracle.sc8_to_fc32(x[s]), decim) for s in range(2)]
    return x, want


@pytest.mark.parametrize("decim", [2, 3, 4, 5, 6, 7, 8, 11, 12, 13, 14, 15, 16, 24, 64])
@pytest.mark.parametrize("fmt", [0, 1, 2])
def test_decimate_kernel_bit_exact(lt, oracle, decim, fmt):
    """rational_resampler_ccc(1, D) for any integer D (examples/cell_search_file.py:50-57): the
    tiled kernel (2, 3, 4, 6, 8, 12), the streaming kernel (16) and the general kernel."""
    rng = np.random.default_rng(decim + 100 * fmt)
    # 1000 outputs: boundary segments only; 5000: whole segments of the streaming kernels too
    for n_out in (1000, 5000):
        x, want = _decim_case(oracle, rng, decim, fmt, n_out * decim)
        got = lt.kernel_decimate(x, decim, fmt)
        for s in range(2):
            assert np.array_equal(got[s].view(np.uint32), want[s].view(np.uint32)), n_out


@pytest.mark.parametrize("decim", [2, 4, 8, 12, 15, 16])
def test_general_decimator_agrees_with_tuned_kernels(lt, oracle, decim):
    """Debug flag 1 (debug build of the library only: lib/libltetrigger_b200_debug.so) routes every rate
    through decimate_any_kernel: three independent kernels and the oracle give the same bits."""
    from ltetrigger_b200 import _abi as A
    D = A.debug_lib()
    assert not hasattr(lt.lib(), "no_such") and "ltb_debug_set_flag" not in A.SYMBOLS
    rng = np.random.default_rng(decim)
    x, want = _decim_case(oracle, rng, decim, 0, 777 * decim)
    tuned = lt.kernel_decimate(x, decim, 0)
    D.ltb_debug_set_flag(1, 1)
    try:
        general = lt.kernel_decimate(x, decim, 0, L=D)
    finally:
        D.ltb_debug_set_flag(1, 0)
    for s in range(2):
        assert np.array_equal(general[s].view(np.uint32), want[s].view(np.uint32))
        assert np.array_equal(tuned[s].view(np.uint32), want[s].view(np.uint32))
    if decim in (4, 8, 12, 15):              # these rates also have the tiled kernel (flag bit 1)
        D.ltb_debug_set_flag(1, 2)
        try:
            tiled = lt.kernel_decimate(x, decim, 0, L=D)
        finally:
            D.ltb_debug_set_flag(1, 0)
        for s in range(2):
            assert np.array_equal(tiled[s].view(np.uint32), want[s].view(np.uint32))


# ---- engine: the four bundled test_frames ------------------------------------------------
@pytest.mark.parametrize("name", list(FIXTURES))
def test_fixture_records_bit_exact(lt, oracle, name):
    x, decim, cell_id = load_fixture(name, 0.5)
    trig = lt.Trigger(n_streams=1, decim=decim, psr_threshold=4.0, max_chunk=96000 * decim)
    got = trig.run(x[None, :])
    want = oracle.trigger_run(x[None, :], decim=decim, psr_threshold=4.0)
    assert_recs_equal(got, want)
    cells = got[(got["flags"] & lt.F_CELL) != 0]
    assert set(cells["cell_id"].tolist()) == {cell_id}
    assert ((cells["flags"] & lt.F_CP_NORM) != 0).all()
    # reference-class FFT evaluation: same decisions, magnitudes within 1e-4 relative
    ref = oracle.trigger_run(x[None, :], decim=decim, psr_threshold=4.0, conv_mode=oracle.CONV_FFT)
    for f in ("win_start", "emit_start", "flags", "peak_pos", "m0", "m1", "n_id_1", "cell_id"):
        assert (got[f] == ref[f]).all(), f
    np.testing.assert_allclose(got["psr"], ref["psr"], rtol=1e-4)
    np.testing.assert_allclose(got["peak_value"], ref["peak_value"], rtol=1e-4)
    # accessors (lib/pss_impl.h:95-100)
    k = cell_id % 3
    st = trig.stats(0, k)
    assert st.tracking == 1 and st.tracking_score == 16.0
    assert st.max_psr == want[want["n_id_2"] == k]["psr"].max()


@pytest.mark.parametrize("name", list(FIXTURES))
def test_fixture_records_fft_correlator(lt, oracle, name):
    """The engine with corr_mode = LTB_CORR_FFT: records bit-identical to the oracle's ORC_CONV_OS
    mode; decisions identical to the direct mode and to the reference's known answers."""
    x, decim, cell_id = load_fixture(name, 0.5)
    trig = lt.Trigger(n_streams=1, decim=decim, psr_threshold=4.0, max_chunk=96000 * decim, corr_mode=lt.CORR_FFT)
    got = trig.run(x[None, :])
    want = oracle.trigger_run(x[None, :], decim=decim, psr_threshold=4.0, conv_mode=oracle.CONV_OS)
    assert_recs_equal(got, want)
    cells = got[(got["flags"] & lt.F_CELL) != 0]
    assert set(cells["cell_id"].tolist()) == {cell_id}
    direct = oracle.trigger_run(x[None, :], decim=decim, psr_threshold=4.0)
    for f in ("win_start", "emit_start", "flags", "peak_pos", "m0", "m1", "n_id_1", "cell_id"):
        assert (got[f] == direct[f]).all(), f
    np.testing.assert_allclose(got["psr"], direct["psr"], rtol=1e-4)


def test_fft_correlator_chunking_and_batch(lt, oracle):
    """Blocks are aligned to absolute sample indices and only whole blocks are evaluated, so ragged
    chunk sizes give the same records; batched noisy streams against the oracle."""
    from ltetrigger_b200 import synth
    iq, ids = synth.batch(6, 384000, 0.0, master_seed=77)
    want = oracle.trigger_run(iq, conv_mode=oracle.CONV_OS)
    for chunk in (8 * 100, 8 * 1117, 8 * 9001, 384000):        # 800 < one 896-sample block: calls without a new block
        trig = lt.Trigger(n_streams=6, decim=1, max_chunk=chunk, corr_mode=lt.CORR_FFT)
        got = trig.run(iq, chunk=chunk)
        assert_recs_equal(got, want)
    trig.reset()
    assert_recs_equal(trig.run(iq), want)


def test_chunking_invariance(lt, oracle):
    x, decim, _ = load_fixture("25prb", 0.4)
    want = oracle.trigger_run(x[None, :], decim=decim)
    for chunk in (32 * 100, 32 * 2999, 32 * 12000):
        trig = lt.Trigger(n_streams=1, decim=decim, max_chunk=chunk)
        got = trig.run(x[None, :], chunk=chunk)
        assert_recs_equal(got, want)


def test_sc16_input(lt, oracle):
    from ltetrigger_b200 import synth
    x = synth.capture(200, 384000, snr_db=8.0, seed=3)
    iq = synth.to_sc16(x)[None]
    trig = lt.Trigger(n_streams=1, decim=1, input_format=lt.FMT_SC16, max_chunk=384000)
    got = trig.run(iq)
    want = oracle.trigger_run(iq, decim=1, fmt=1)
    assert_recs_equal(got, want)
    assert 200 in got["cell_id"]


@pytest.mark.parametrize("decim", [1, 16])
def test_sc8_input(lt, oracle, decim):
    from ltetrigger_b200 import synth
    x = synth.capture(77, 288000 * decim, snr_db=8.0, seed=5, decim=decim)
    iq = synth.to_sc8(x)[None]
    trig = lt.Trigger(n_streams=1, decim=decim, input_format=lt.FMT_SC8, max_chunk=96000 * decim)
    got = trig.run(iq, chunk=96000 * decim)
    want = oracle.trigger_run(iq, decim=decim, fmt=2)
    assert_recs_equal(got, want)
    assert 77 in got["cell_id"]


@pytest.mark.parametrize("decim,fmt", [(12, 0), (12, 1), (3, 0), (6, 2), (5, 0), (20, 1)])
def test_any_integer_sample_rate(lt, oracle, decim, fmt):
    """Sample rates that are not 1.92 MHz * 2^k (23.04 Msps = 15 MHz LTE at D = 12, ...), fed in
    ragged chunks: records identical to the oracle's."""
    from ltetrigger_b200 import synth
    x = np.stack([synth.capture(c, 240000 * decim, snr_db=7.0, decim=decim, seed=c) for c in (91, 502)])
    iq = x if fmt == 0 else (synth.to_sc16(x) if fmt == 1 else synth.to_sc8(x))
    chunk = 8 * decim * 7001
    trig = lt.Trigger(n_streams=2, decim=decim, input_format=fmt, max_chunk=chunk)
    got = trig.run(iq, chunk=chunk)
    want = oracle.trigger_run(iq, decim=decim, fmt=fmt)
    assert_recs_equal(got, want)
    assert {91, 502} <= set(got["cell_id"].tolist())


def test_extended_cp_capture(lt, oracle):
    """Extended-CP cells: srslte_sync_detect_cp picks the extended hypothesis and the SSS symbol
    is taken 128 + 32 samples before the PSS (lib/sss_impl.cc:104-110)."""
    from ltetrigger_b200 import synth
    x = np.stack([synth.capture(c, 384000, snr_db=9.0, seed=c, ext_cp=e) for c, e in ((311, True), (40, False), (167, True))])
    trig = lt.Trigger(n_streams=3, decim=1, max_chunk=384000)
    got = trig.run(x)
    want = oracle.trigger_run(x)
    assert_recs_equal(got, want)
    for s, (c, e) in enumerate(((311, True), (40, False), (167, True))):
        cells = got[(got["stream"] == s) & ((got["flags"] & lt.F_CELL) != 0)]
        assert len(cells) > 5 and set(cells["cell_id"].tolist()) == {c}
        assert (((cells["flags"] & lt.F_CP_NORM) != 0) == (not e)).all()


@pytest.mark.parametrize("corr", ["direct", "fft"])
def test_tdd_sss_position(lt, oracle, corr):
    """Frame structure type 2 (not in the reference): SSS three symbols before the PSS, normal and
    extended CP.  Engine records against the oracle's TDD restatement, and the FDD setting on the same
    capture does not find the cell."""
    from ltetrigger_b200 import synth
    cases = ((301, False), (17, True), (440, False))
    x = np.stack([synth.capture(c, 384000, snr_db=9.0, seed=c, ext_cp=e, tdd=True) for c, e in cases])
    mode, conv = (lt.CORR_FFT, oracle.CONV_OS) if corr == "fft" else (lt.CORR_DIRECT, oracle.CONV_DIRECT)
    trig = lt.Trigger(n_streams=3, decim=1, max_chunk=384000, corr_mode=mode, frame_type=lt.FRAME_TDD)
    got = trig.run(x)
    want = oracle.trigger_run(x, conv_mode=conv | oracle.FRAME_TDD)
    assert_recs_equal(got, want)
    for s, (c, e) in enumerate(cases):
        cells = got[(got["stream"] == s) & ((got["flags"] & lt.F_CELL) != 0)]
        assert len(cells) > 5 and set(cells["cell_id"].tolist()) == {c}
        assert (((cells["flags"] & lt.F_CP_NORM) != 0) == (not e)).all()
    fdd = lt.Trigger(n_streams=3, decim=1, max_chunk=384000, corr_mode=mode).run(x)
    assert_recs_equal(fdd, oracle.trigger_run(x, conv_mode=conv))
    for s, (c, e) in enumerate(cases):
        ids = fdd[(fdd["stream"] == s) & ((fdd["flags"] & lt.F_CELL) != 0)]["cell_id"]
        assert (ids != c).all()


def test_synthetic_snr_sweep_batched(lt, oracle):
    """Reduced config C4: 16 streams x 0.25 s per SNR point; event lists identical to the oracle's."""
    from ltetrigger_b200 import synth
    n = 480000
    det = {}
    for snr in (-10.0, -4.0, 0.0, 10.0):
        iq, ids = synth.batch(16, n, snr, master_seed=100 + int(snr))
        trig = lt.Trigger(n_streams=16, decim=1, psr_threshold=4.0, max_chunk=160000)
        got = trig.run(iq, chunk=160000)
        want = oracle.trigger_run(iq, decim=1, psr_threshold=4.0)
        assert_recs_equal(got, want)
        ok = 0
        for s in range(16):
            c = got[(got["stream"] == s) & ((got["flags"] & lt.F_CELL) != 0)]["cell_id"]
            ok += int(len(c) > 0 and np.bincount(c).argmax() == ids[s])
        det[snr] = ok
    assert det[10.0] == 16 and det[0.0] >= 14


def test_decimated_synthetic_with_cfo(lt, oracle):
    from ltetrigger_b200 import synth
    x = np.stack([synth.capture(c, 16 * 300000, snr_db=6.0, decim=16, seed=c, cfo_hz=f)
                  for c, f in ((369, 400.0), (12, -700.0))])
    trig = lt.Trigger(n_streams=2, decim=16, max_chunk=16 * 100000)
    got = trig.run(x, chunk=16 * 100000)
    want = oracle.trigger_run(x, decim=16)
    assert_recs_equal(got, want)
    assert {369, 12} <= set(got["cell_id"].tolist())


def test_edge_cases(lt, oracle):
    trig = lt.Trigger(n_streams=2, decim=1, max_chunk=96000)
    # shorter than the lookahead: no call; all zeros: NaN PSR, nothing emitted
    assert len(trig.process(np.zeros((2, 18360), np.complex64))) == 0
    recs = trig.process(np.zeros((2, 96000), np.complex64))
    assert np.isnan(recs["psr"]).all() and (recs["flags"] & lt.F_EMIT == 0).all()
    with pytest.raises(lt.LtbError):
        trig.process(np.zeros((2, 1001), np.complex64))        # not a multiple of 8
    # reset returns to the constructed state
    x, _, _ = load_fixture("6prb", 0.2)
    iq = np.stack([x, x])
    trig.reset()
    a = trig.run(iq).copy()
    trig.reset()
    b = trig.run(iq)
    assert a.tobytes() == b.tobytes()
    assert_recs_equal(a, oracle.trigger_run(iq))
    # threshold setter with clamp (python/downlink_trigger_c.py:63-73)
    trig.reset()
    trig.set_psr_threshold(0.1)
    assert trig.stats(0, 0).psr_threshold == 1.5
    c = trig.run(iq)
    assert_recs_equal(c, oracle.trigger_run(iq, psr_threshold=1.5))


def test_record_all_off_keeps_only_emitted(lt):
    x, _, _ = load_fixture("6prb", 0.2)
    full = lt.Trigger(n_streams=1, max_chunk=384000).run(x[None, :])
    emit = lt.Trigger(n_streams=1, max_chunk=384000, record_all=False).run(x[None, :])
    assert_recs_equal(emit, full[(full["flags"] & lt.F_EMIT) != 0])


def test_large_cfo_exercises_phasor_scan(lt, oracle):
    """Large carrier offsets make srslte_cfo_correct's phase wrap often and visit many float
    binades: the segmented scan in the track kernel must stay bit-identical to the oracle's
    sample-by-sample loop, also over full half-frames (keep_halfframes)."""
    from ltetrigger_b200 import synth
    cfos = (-6000.0, 4000.0, 7400.0, -250.0, 15.0, 0.0)
    x = np.stack([synth.capture(30 + 7 * i, 480000, snr_db=15.0, seed=i, cfo_hz=f) for i, f in enumerate(cfos)])
    trig = lt.Trigger(n_streams=len(cfos), decim=1, max_chunk=480000, psr_threshold=2.5, keep_halfframes=True)
    got = trig.run(x)
    want = oracle.trigger_run(x, decim=1, psr_threshold=2.5)
    assert_recs_equal(got, want)
    assert ((got["flags"] & lt.F_TRACKING) != 0).sum() > 100
    # emitted half-frames (CFO-corrected when tracking) against the oracle's pss block output
    n_emit = int(((got["flags"] & lt.F_EMIT) != 0).sum())
    hfs = trig.fetch_halfframes(n_emit)
    k = 0
    for s in range(len(cfos)):
        buf = np.concatenate([np.zeros(960, np.complex64), x[s]])
        for r in range(3):
            blk = oracle.Pss(r, 2.5)
            pos = 960
            while pos - 960 + oracle.LOOKAHEAD <= x.shape[1]:
                nout, ncons, out, rec = blk.work(buf, pos)
                if nout:
                    assert np.array_equal(hfs[k].view(np.uint32), out.view(np.uint32)), (s, r, pos)
                    k += 1
                pos += ncons
    assert k == n_emit


def test_two_calls_in_flight_match_synchronous_calls(lt, oracle):
    """submit/submit/collect/collect (two calls in flight) gives the same records as the
    synchronous path; a third submit and a process call while calls are pending are refused."""
    import torch
    x, decim, _ = load_fixture("25prb", 0.3)
    n = len(x) // 3 // (8 * decim) * (8 * decim)
    want = oracle.trigger_run(x[None, :3 * n], decim=decim)
    d = torch.from_numpy(x[:3 * n].copy()).cuda()
    trig = lt.Trigger(n_streams=1, decim=decim, max_chunk=n)
    trig.submit_device_ptr(d.data_ptr(), 0, n)
    trig.submit_device_ptr(d.data_ptr() + 8 * n, 0, n)
    with pytest.raises(lt.LtbError):
        trig.submit_device_ptr(d.data_ptr() + 16 * n, 0, n)
    with pytest.raises(lt.LtbError):
        trig.process_device_ptr(d.data_ptr() + 16 * n, 0, n)
    got = [trig.collect().copy()]
    trig.submit_device_ptr(d.data_ptr() + 16 * n, 0, n)
    got += [trig.collect().copy(), trig.collect().copy()]
    with pytest.raises(lt.LtbError):
        trig.collect()
    recs = np.concatenate(got)
    recs = recs[np.lexsort((recs["win_index"], recs["n_id_2"], recs["stream"]))]
    assert_recs_equal(recs, want)


@pytest.mark.parametrize("pipe", ["overlap", "serial"])
def test_pipeline_modes_give_the_same_records(lt, oracle, pipe):
    """The two ways the kernels of consecutive calls are scheduled (track + SSS of call i under the front end of call
    i+1, or everything on one stream) only move kernels in time: with two calls in flight over five chunks every
    record equals the oracle's."""
    import torch
    x, decim, _ = load_fixture("100prb", 0.3)
    n = len(x) // 5 // (8 * decim) * (8 * decim)
    iq = np.stack([x[:5 * n], np.roll(x[:5 * n], 12345)])
    want = oracle.trigger_run(iq, decim=decim, conv_mode=oracle.CONV_OS)
    d = torch.from_numpy(iq.copy()).cuda()
    mode = {"overlap": lt.PIPE_OVERLAP, "serial": lt.PIPE_SERIAL}[pipe]
    trig = lt.Trigger(n_streams=2, decim=decim, max_chunk=n, corr_mode=lt.CORR_FFT, pipeline=mode)
    got = []
    trig.submit_device_ptr(d.data_ptr(), 8 * 5 * n, n)
    for k in range(1, 5):
        trig.submit_device_ptr(d.data_ptr() + 8 * k * n, 8 * 5 * n, n)
        got.append(trig.collect().copy())
    got.append(trig.collect().copy())
    trig.close()
    recs = np.concatenate(got)
    recs = recs[np.lexsort((recs["win_index"], recs["n_id_2"], recs["stream"]))]
    assert_recs_equal(recs, want)


def test_two_host_calls_in_flight(lt, oracle):
    """ltb_trigger_submit_host: the copy of call i+1 overlaps call i; records equal the oracle's."""
    import torch
    x, decim, _ = load_fixture("50prb", 0.3)
    n = len(x) // 3 // (8 * decim) * (8 * decim)
    want = oracle.trigger_run(x[None, :3 * n], decim=decim)
    h = torch.from_numpy(x[:3 * n].copy()).pin_memory()
    trig = lt.Trigger(n_streams=1, decim=decim, max_chunk=n)
    got = []
    trig.submit_host_ptr(h.data_ptr(), 0, n)
    trig.submit_host_ptr(h.data_ptr() + 8 * n, 0, n)
    got.append(trig.collect().copy())
    trig.submit_host_ptr(h.data_ptr() + 16 * n, 0, n)
    got += [trig.collect().copy(), trig.collect().copy()]
    recs = np.concatenate(got)
    recs = recs[np.lexsort((recs["win_index"], recs["n_id_2"], recs["stream"]))]
    assert_recs_equal(recs, want)


@pytest.mark.parametrize("corr", ["direct", "fft"])
def test_full_path_scaling_and_permutation_properties(lt, corr):
    """Size-independent properties of the whole path at a bench-like size (no oracle: 48 streams x
    30.72 Msps x 100 ms, decimate by 16): doubling the input is exact in float32, so every
    decision, index, PSR and CFO keeps its bits while powers scale by 4 (SSS correlations by 2);
    and streams are independent, so permuting them permutes the record lists."""
    from ltetrigger_b200 import synth
    rng = np.random.default_rng(11)
    n, S = 16 * 192000, 48
    base = np.stack([synth.capture(int(c), n, snr_db=4.0, decim=16, seed=int(c)) for c in rng.integers(0, 504, 6)])
    x = np.empty((S, n), np.complex64)
    for s in range(S):
        x[s] = np.roll(base[s % 6], int(rng.integers(0, n))) + 0.05 * rand_c64(rng, n)
    mode = lt.CORR_FFT if corr == "fft" else lt.CORR_DIRECT

    def run(iq):
        trig = lt.Trigger(n_streams=S, decim=16, max_chunk=n, corr_mode=mode)
        out = trig.run(iq)
        trig.close()
        return out

    a, b = run(x), run(2 * x)
    assert len(a) == len(b) and len(a) > 20 * S
    for f in a.dtype.names:
        if f in ("peak_value",):
            assert np.array_equal(4 * a[f], b[f]), f
        elif f in ("m0_val", "m1_val"):
            assert np.array_equal(2 * a[f], b[f]) or np.array_equal(4 * a[f], b[f]), f
        else:
            assert a[f].tobytes() == b[f].tobytes(), f
    perm = rng.permutation(S)
    c = run(x[perm])
    for s_new, s_old in enumerate(perm):
        ra, rc = a[a["stream"] == s_old], c[c["stream"] == s_new].copy()
        rc["stream"] = s_old
        assert ra.tobytes() == rc.tobytes(), (s_new, s_old)


def test_c_abi_error_behaviour(lt):
    """Return codes of the engine entry points (srsLTE convention 0 / -1 / -2, lib/sss_impl.cc:119): a
    record buffer that is too small is reported with the needed count and leaves the call pending (collect
    again with a larger buffer: nothing is lost), oversized and misaligned chunks are refused without side
    effects, optional features say so when they are off."""
    import ctypes as C
    from ltetrigger_b200 import _abi as A
    L = lt.lib()
    x, _, _ = load_fixture("6prb", 0.2)
    trig = lt.Trigger(n_streams=1, max_chunk=96000)
    full = trig.run(x[None, :96000 * 4]).copy()
    trig.reset()
    small = np.zeros(5, A.WINDOW_REC)
    n = C.c_int32(0)
    iq = np.ascontiguousarray(x[:96000])
    rc = L.ltb_trigger_process_host(trig._h, iq.ctypes.data, 0, 96000, small.ctypes.data, 5, C.byref(n))
    assert rc == lt.ERROR_INVALID_INPUTS and n.value > 5              # says how many there are
    n_first = n.value
    # the call is still pending: a new submit is refused, a retry with enough room gets every record
    assert L.ltb_trigger_submit_host(trig._h, iq.ctypes.data, 0, 96000) == lt.SUCCESS      # second slot
    assert L.ltb_trigger_submit_host(trig._h, iq.ctypes.data, 0, 96000) == lt.ERROR_INVALID_INPUTS
    retry = np.zeros(n_first, A.WINDOW_REC)
    assert L.ltb_trigger_collect(trig._h, retry.ctypes.data, n_first, C.byref(n)) == lt.SUCCESS and n.value == n_first
    first = full[full["win_start"] + lt.LOOKAHEAD <= 96000]          # the calls the first chunk allows, in record order
    assert len(first) == n_first and retry.tobytes() == first.tobytes()
    second = np.zeros(64, A.WINDOW_REC)
    assert L.ltb_trigger_collect(trig._h, second.ctypes.data, 64, C.byref(n)) == lt.SUCCESS
    n_second = n.value
    big = np.zeros(96008, np.complex64)
    assert L.ltb_trigger_process_host(trig._h, big.ctypes.data, 0, 96008, small.ctypes.data, 5, C.byref(n)) == lt.ERROR_INVALID_INPUTS
    assert L.ltb_trigger_process_host(trig._h, None, 0, 96000, small.ctypes.data, 5, C.byref(n)) == lt.ERROR_INVALID_INPUTS
    assert L.ltb_trigger_collect(trig._h, small.ctypes.data, 5, C.byref(n)) == lt.ERROR_INVALID_INPUTS   # nothing submitted
    with pytest.raises(lt.LtbError):
        trig.fetch_halfframes(4)                                      # keep_halfframes is off
    st = A.PssStats()
    assert L.ltb_trigger_get_stats(trig._h, 3, 0, C.byref(st)) == lt.ERROR_INVALID_INPUTS
    assert L.ltb_trigger_get_stats(trig._h, 0, 3, C.byref(st)) == lt.ERROR_INVALID_INPUTS
    # the refused calls left the engine where it was: the remaining chunks give the remaining records
    rest = trig.run(x[None, 96000 * 2:96000 * 4])
    assert len(rest) == len(full) - n_first - n_second
    for k in range(3):
        fk, rk = full[full["n_id_2"] == k], rest[rest["n_id_2"] == k]
        assert fk[len(fk) - len(rk):].tobytes() == rk.tobytes()
    with pytest.raises(lt.LtbError):
        lt.Trigger(n_streams=1, device=99)                            # no such CUDA device


def test_fft_correlator_edge_cases(lt, oracle):
    """LTB_CORR_FFT on degenerate input: shorter than the lookahead -> no call; all zeros -> NaN PSR and
    nothing emitted; a stream that starts mid-frame and a silent stream next to it; reset."""
    trig = lt.Trigger(n_streams=2, decim=1, max_chunk=96000, corr_mode=lt.CORR_FFT)
    assert len(trig.process(np.zeros((2, 18360), np.complex64))) == 0
    recs = trig.process(np.zeros((2, 96000), np.complex64))
    assert len(recs) > 0 and np.isnan(recs["psr"]).all() and (recs["flags"] & lt.F_EMIT == 0).all()
    x, _, _ = load_fixture("6prb", 0.3)
    iq = np.stack([x[5000:5000 + 480000], np.zeros(480000, np.complex64)])
    trig.reset()
    a = trig.run(iq).copy()
    want = oracle.trigger_run(iq, conv_mode=oracle.CONV_OS)
    assert_recs_equal(a, want)
    assert 123 in a[a["stream"] == 0]["cell_id"] and np.isnan(a[a["stream"] == 1]["psr"]).all()
    trig.reset()
    assert trig.run(iq).tobytes() == a.tobytes()


def test_two_engines_from_two_host_threads(lt, oracle):
    """The ABI's threading contract: one host thread per object at a time, different objects from
    different threads concurrently (the reference runs one scheduler thread per block).  Two engines
    with different rates, formats and correlators run in parallel threads and both match the oracle."""
    import threading
    from ltetrigger_b200 import synth
    xa = np.stack([synth.capture(c, 16 * 200000, snr_db=6.0, decim=16, seed=c) for c in (21, 400)])
    xb16 = synth.to_sc16(np.stack([synth.capture(c, 480000, snr_db=3.0, seed=c) for c in (77, 78, 79)]))
    out, err = {}, []

    def run(key, iq, **kw):
        try:
            trig = lt.Trigger(n_streams=iq.shape[0], **kw)
            parts = []
            for _ in range(3):                       # several passes so the two threads really overlap
                trig.reset()
                parts.append(trig.run(iq, chunk=kw["max_chunk"]).copy())
            assert parts[0].tobytes() == parts[1].tobytes() == parts[2].tobytes()
            out[key] = parts[0]
            trig.close()
        except Exception as e:                       # surfaced in the main thread below
            err.append((key, e))

    ta = threading.Thread(target=run, args=("a", xa), kwargs=dict(decim=16, max_chunk=16 * 50000, corr_mode=lt.CORR_FFT))
    tb = threading.Thread(target=run, args=("b", xb16), kwargs=dict(decim=1, max_chunk=96000, input_format=lt.FMT_SC16))
    ta.start(); tb.start(); ta.join(); tb.join()
    assert not err, err
    assert_recs_equal(out["a"], oracle.trigger_run(xa, decim=16, conv_mode=oracle.CONV_OS))
    assert_recs_equal(out["b"], oracle.trigger_run(xb16, decim=1, fmt=1))


@pytest.mark.parametrize("decim", [16, 12, 8, 1])
def test_device_input_that_is_only_sample_aligned(lt, oracle, decim):
    """The streaming decimators copy 16-byte aligned segments; a device buffer whose base or row
    stride is merely sample aligned (8 bytes for fc32) runs a slower kernel and gives the same
    records.  A pointer that is not even sample aligned is refused."""
    import torch
    from ltetrigger_b200 import synth
    n = 8 * decim * 12000
    x = np.stack([synth.capture(c, n, snr_db=8.0, decim=decim, seed=c) for c in (55, 56)])
    want = oracle.trigger_run(x, decim=decim)
    buf = torch.zeros(2 * (n + 3) + 1, dtype=torch.complex64, device="cuda")
    view = buf[1:1 + 2 * (n + 3)].view(2, n + 3)          # base 8 bytes off, row stride 8 (n + 3) bytes
    view[:, :n] = torch.from_numpy(x).cuda()
    assert view.data_ptr() % 16 == 8
    trig = lt.Trigger(n_streams=2, decim=decim, max_chunk=n)
    got = trig.process_device_ptr(view.data_ptr(), 8 * (n + 3), n).copy()
    got = got[np.lexsort((got["win_index"], got["n_id_2"], got["stream"]))]
    assert_recs_equal(got, want)
    with pytest.raises(lt.LtbError):
        trig.process_device_ptr(view.data_ptr() + 4, 8 * (n + 3), n)


def test_product_reports_the_transmitted_truth_without_the_oracle(lt):
    """The CUDA path against what was transmitted (tests/test_oracle_truth.py does the same for the oracle): known cell id,
    CP type, frame timing and carrier offset of seeded synthetic captures, at the search rate and through both decimating
    front ends at D = 16 (group delay (525 - 1) / 2 input samples)."""
    from ltetrigger_b200 import synth
    rng = np.random.default_rng(4242)
    for decim, frontend, fs in ((1, lt.FRONTEND_FP32, 0.0), (16, lt.FRONTEND_FP32, 0.0), (16, lt.FRONTEND_TC_INT, 8.0)):
        cases = [(int(rng.integers(0, 504)), int(rng.integers(0, 19200 * decim)), float(rng.uniform(-2500.0, 2500.0)),
                  bool(i == 2)) for i in range(6)]
        # the delayed frame boundary on a whole search-rate sample (a fractional one biases the half-symbol CFO estimate:
        # a Zadoff-Chu sequence sampled off the grid looks frequency-shifted)
        cases = [(c, o + (262 - o) % decim, f, e) for c, o, f, e in cases]
        n = 19200 * decim * 24
        x = np.stack([synth.capture(c, n, snr_db=12.0, decim=decim, seed=50 + i, offset=o, cfo_hz=f, ext_cp=e)
                      for i, (c, o, f, e) in enumerate(cases)])
        trig = lt.Trigger(n_streams=len(cases), decim=decim, psr_threshold=4.0, max_chunk=96000 * decim,
                          corr_mode=lt.CORR_FFT, frontend_mode=frontend, fc32_full_scale=fs)
        recs = trig.run(x, chunk=96000 * decim)
        trig.close()
        delay = (525 - 1) / 2.0 / decim if decim == 16 else 0.0
        for s, (cell, offset, cfo_hz, ext_cp) in enumerate(cases):
            t = recs[(recs["stream"] == s) & ((recs["flags"] & lt.F_CELL) != 0)]
            what = (decim, frontend, cell, offset, cfo_hz, ext_cp)
            assert len(t) >= 4 and set(t["cell_id"].tolist()) == {cell}, what
            assert set(((t["flags"] & lt.F_CP_NORM) != 0).tolist()) == {not ext_cp}, what
            truth = (-offset / decim + delay) % 9600.0
            d = (t["emit_start"] % 9600 - truth + 4800.0) % 9600.0 - 4800.0
            assert np.abs(d).max() == 0.0, (what, truth)
            trk = t[(t["flags"] & lt.F_TRACKING) != 0]
            assert len(trk) and abs(float(trk["mean_cfo"][-1]) - cfo_hz / 15000.0) < 0.01, what
