"""LTB_FRONTEND_TC_INT on the CPU: the tensor-core kernel's k-step schedule (csrc/ltb_tc_frontend.cuh) replayed
in numpy from the product's own tap table -- byte rows times the [208 x 32-byte] digit table, accumulated at the
column offset of each k-step, the four weights recombined, the q = 0..3 diagonal summed -- equals the oracle's
int64 restatement bit for bit, for every input format.  This pins the table layout and the arithmetic
contract without a GPU; tests/test_gpu_tc.py then compares the kernel itself with the same oracle."""
import numpy as np
import pytest


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _rows_as_bytes(comp, fmt, q_inv=None, decim=16):
    """One component of one stream -> [n_rows, K bytes] as the transform warps write them (tc_split); a row is 16
    outputs = 16 decim samples."""
    if fmt == 1:      # sc16: lo byte flipped to signed, hi byte signed
        u = comp.astype(np.int16).view(np.uint16)
        lo = ((u & 0xff) ^ 0x80).astype(np.uint8).view(np.int8).astype(np.int64)
        hi = (u >> 8).astype(np.uint8).view(np.int8).astype(np.int64)
        k = np.stack([lo, hi], axis=-1).reshape(-1, 32 * decim)
    elif fmt == 2:    # sc8: signed bytes as they are
        k = comp.astype(np.int64).reshape(-1, 16 * decim)
    else:             # fc32: two fused multiply-adds, unsigned bytes, byte 3 is whatever the float holds
        x = comp.astype(np.float32).astype(np.float64)
        u = np.float32(x * np.float64(q_inv) + 0.5)                       # float64 product of two float32 is exact; one rounding = fma
        u = np.where(np.isnan(u), np.float32(0), np.clip(u, np.float32(0), np.float32(1))).astype(np.float32)
        t = np.float32(u.astype(np.float64) * 8388606.0 + 8388609.0)
        w = t.view(np.uint32)
        k = np.stack([(w >> (8 * b)) & 0xff for b in range(4)], axis=-1).astype(np.int64).reshape(-1, 64 * decim)
    return k


def _ntaps(decim):
    n = int((7.0 / 0.1102 + 8.7) / (22.0 * 0.1 / decim))
    return n + 1 - (n & 1)


def _replay(comp, fmt, btab, sum_t, q_inv=None, out_scale=None, decim=16):
    """float32 outputs of one component, computed the way the kernel does (three halo rows of zero samples
    before the stream, as the engine's zeroed history)."""
    import math
    comp = np.concatenate([np.zeros(3 * 16 * decim, comp.dtype), comp])
    rows = _rows_as_bytes(comp, fmt, q_inv, decim)
    n_rows, kbytes = rows.shape
    ksteps = kbytes // 32
    spk = {0: 8, 1: 16, 2: 32}[fmt]                                        # samples per k-step
    g = math.gcd(spk, decim)
    nd = (_ntaps(decim) - 1 + (decim - g) + spk - 1) // decim + 1
    nstep = (4 * nd + 15) // 16 * 16
    acc = np.zeros((n_rows, 208), np.int64)
    for s in range(ksteps):
        a = rows[:, 32 * s:32 * s + 32]
        off = s * spk
        u0, ph = off // decim, (off % decim) // g
        b, col = btab[:, 32 * ph:32 * ph + 32], 4 * u0
        n = 208 if s == 0 else nstep
        assert col + n <= 208
        part = a @ b[:n].astype(np.int64).T
        assert np.abs(part).max() < 2 ** 31
        acc[:, col:col + n] += part
    assert np.abs(acc).max() < 2 ** 31                                    # int32 accumulators in TMEM
    val = acc[:, 0::4] + 256 * acc[:, 1::4] + 65536 * acc[:, 2::4] + 16777216 * acc[:, 3::4]    # [n_rows, 52] column groups u
    assert not val[:, 49:].any()
    out = np.zeros((n_rows, 16), np.int64)
    for q in range(4):                                                    # u = 16 q + r: row b feeds output r of row b + q
        nu = 16 if q < 3 else 1
        out[q:, :nu] += val[:n_rows - q if q else n_rows, 16 * q:16 * q + nu]
    shift = 23 + int(math.log2(decim))
    c = {0: -16384 * sum_t, 1: 128 * sum_t, 2: 0}[fmt]
    sc = {0: out_scale, 1: np.float32(2.0 ** -(shift + 15)), 2: np.float32(2.0 ** -(shift + 7))}[fmt]
    return ((out.reshape(-1) + c).astype(np.float32) * np.float32(sc))[48:]   # int64 -> float32 rounds to nearest even, as I2F.S64 does


@pytest.mark.parametrize("fmt,decim", [(0, 16), (1, 16), (2, 16), (0, 8), (1, 8), (2, 8), (0, 4), (1, 4), (0, 2), (0, 12), (1, 12),
                                       (0, 24), (1, 24), (2, 24), (0, 32), (1, 32), (2, 32)])
def test_kernel_schedule_equals_oracle(oracle, fmt, decim):
    import math
    import ltetrigger_b200 as lt
    btab, sum_t = lt.tables.tc_btab(fmt, decim)
    rng = np.random.default_rng(10 + fmt + 100 * decim)
    n = 16 * decim * 12
    if fmt == 0:
        fs = 2.5
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        x[100:108] = [np.nan, np.inf, -np.inf, fs, -fs, 7 * fs, 1e-40, -0.0]
        q_inv = np.float32(0.5 / fs)
        shift = 23 + int(math.log2(decim))
        out_scale = np.float32(fs / 4194303.0 * 2.0 ** (8 - shift))
        want = oracle.decimate_tcint_fc32(x, fs, decim)
        got = _replay(x.real, 0, btab, sum_t, q_inv, out_scale, decim) + 1j * _replay(x.imag, 0, btab, sum_t, q_inv, out_scale, decim)
    else:
        lo, hi, dt = (-32768, 32767, np.int16) if fmt == 1 else (-128, 127, np.int8)
        iq = rng.integers(lo, hi + 1, size=(n, 2)).astype(dt)
        iq[:n // 10] = hi
        iq[n // 10:n // 5] = lo
        want = (oracle.decimate_tcint_sc16 if fmt == 1 else oracle.decimate_tcint_sc8)(iq, decim)
        got = _replay(iq[:, 0], fmt, btab, sum_t, decim=decim) + 1j * _replay(iq[:, 1], fmt, btab, sum_t, decim=decim)
    got = got.astype(np.complex64)
    assert np.array_equal(_bits(got), _bits(want)), int(np.argmax(_bits(got) != _bits(want)))


def test_fixed_point_front_end_is_close_to_float32(oracle):
    """The fc32 fixed-point grid against the canonical float32 decimator: ~1e-6 of the output's rms when the
    declared range is 8 x the signal's rms (the north_star tolerance is 1e-4)."""
    rng = np.random.default_rng(3)
    n = 16 * 3000
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    a = oracle.decimate_tcint_fc32(x, 8.0 * np.sqrt(2.0))
    b = oracle.decimate(x, 16)
    assert np.abs(a - b).max() < 2e-5 * np.sqrt(np.mean(np.abs(b) ** 2))
