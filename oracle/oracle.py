"""ctypes binding of the CPU oracle (oracle/ltetrigger_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libltetrigger_oracle.so")

CONV_DIRECT, CONV_FFT, CONV_OS = 0, 1, 2
FRAME_TDD = 0x100          # or-ed into conv_mode: TDD SSS position
FRONT_TCINT = 0x200        # or-ed into conv_mode of trigger_run: exact integer front end (sc16, decim 16)
OS_STEP = 896
SLOT, HALF, SYM, CONV_LEN, LOOKAHEAD = 960, 9600, 128, 9726, 18365

F_SEARCHED, F_OVER, F_EMIT, F_TRACKING, F_TAG_LOST, F_SSS, F_CELL, F_CP_NORM = (
    0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80)

# numpy mirror of orc_rec (and of ltb_window_rec in include/ltetrigger_b200.h)
REC_DTYPE = np.dtype([
    ("win_start", "<i8"), ("emit_start", "<i8"), ("stream", "<i4"), ("n_id_2", "<i4"),
    ("win_index", "<i4"), ("flags", "<u4"), ("peak_pos", "<i4"), ("score", "<i4"),
    ("psr", "<f4"), ("peak_value", "<f4"), ("cfo", "<f4"), ("mean_cfo", "<f4"),
    ("m0", "<i4"), ("m1", "<i4"), ("m0_val", "<f4"), ("m1_val", "<f4"),
    ("n_id_1", "<i4"), ("cell_id", "<i4"), ("cp_norm_avg", "<f4"), ("cp_ext_avg", "<f4")],
    align=True)
assert REC_DTYPE.itemsize == 88, REC_DTYPE.itemsize


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "ltetrigger_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "CC=" + os.environ.get("ORACLE_CC", "gcc")],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        fp, ip, vp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p
        L.orc_pss_taps.argtypes = [C.c_int, fp, fp]
        L.orc_decim_taps.argtypes = [C.c_int, fp, C.c_int]
        L.orc_decim_taps.restype = C.c_int
        L.orc_sss_tables.argtypes = [C.c_int, ip, ip, ip, ip, ip]
        L.orc_cexptab.argtypes = [fp, fp]
        L.orc_fft128_twiddles.argtypes = [fp, fp]
        L.orc_decimate.argtypes = [vp, C.c_int64, C.c_int, vp]
        L.orc_decimate.restype = C.c_int64
        L.orc_decimate_fast.argtypes = [vp, C.c_int64, C.c_int, vp]
        L.orc_decimate_fast.restype = C.c_int64
        L.orc_decimate_tcint_sc16.argtypes = [vp, C.c_int64, vp]
        L.orc_decimate_tcint_sc16.restype = C.c_int64
        L.orc_decimate_tcint_sc8.argtypes = [vp, C.c_int64, vp]
        L.orc_decimate_tcint_sc8.restype = C.c_int64
        L.orc_decimate_tcint_fc32.argtypes = [vp, C.c_int64, C.c_float, vp]
        L.orc_decimate_tcint_fc32.restype = C.c_int64
        L.orc_decimate_tcint_sc16_d.argtypes = [vp, C.c_int64, C.c_int, vp]
        L.orc_decimate_tcint_sc16_d.restype = C.c_int64
        L.orc_decimate_tcint_sc8_d.argtypes = [vp, C.c_int64, C.c_int, vp]
        L.orc_decimate_tcint_sc8_d.restype = C.c_int64
        L.orc_decimate_tcint_fc32_d.argtypes = [vp, C.c_int64, C.c_int, C.c_float, vp]
        L.orc_decimate_tcint_fc32_d.restype = C.c_int64
        L.orc_trigger_run2.argtypes = [vp, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                       C.c_int, C.c_float, C.c_int, vp, C.c_int]
        L.orc_trigger_run2.restype = C.c_int
        L.orc_sc16_to_fc32.argtypes = [vp, C.c_int64, C.c_float, vp]
        L.orc_sc8_to_fc32.argtypes = [vp, C.c_int64, C.c_float, vp]
        L.orc_pss_corr_window.argtypes = [vp, C.c_int, C.c_int, fp]
        L.orc_pss_corr_stream.argtypes = [vp, C.c_int64, C.c_int, fp]
        L.orc_fft128.argtypes = [vp, vp]
        L.orc_fft1024.argtypes = [vp, vp, C.c_int]
        L.orc_fft1024_twiddles.argtypes = [fp, fp]
        L.orc_os_filter.argtypes = [C.c_int, fp, fp]
        L.orc_pss_corr_os.argtypes = [vp, C.c_int64, vp, vp, vp]
        L.orc_pss_corr_os.restype = C.c_int64
        L.orc_pss_new.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orc_pss_new.restype = vp
        L.orc_pss_free.argtypes = [vp]
        L.orc_pss_work.argtypes = [vp, vp, vp, ip, vp]
        L.orc_pss_work.restype = C.c_int
        for name in ("max_psr", "mean_psr", "mean_cfo", "psr_threshold", "tracking_score"):
            f = getattr(L, "orc_pss_" + name)
            f.argtypes = [vp]
            f.restype = C.c_float
        L.orc_pss_set_psr_threshold.argtypes = [vp, C.c_float]
        L.orc_sss_new.argtypes = [C.c_int]
        L.orc_sss_new.restype = vp
        L.orc_sss_free.argtypes = [vp]
        L.orc_sss_set_frame_type.argtypes = [vp, C.c_int]
        L.orc_sss_work.argtypes = [vp, vp, C.c_int, vp, vp]
        L.orc_sss_work.restype = C.c_int
        L.orc_chain_run.argtypes = [vp, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, vp, C.c_int]
        L.orc_chain_run.restype = C.c_int
        L.orc_trigger_run.argtypes = [vp, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                      C.c_int, C.c_int, vp, C.c_int]
        L.orc_trigger_run.restype = C.c_int
        _lib = L
    return _lib


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def pss_taps(n_id_2):
    hr, hi = np.zeros(128, np.float32), np.zeros(128, np.float32)
    lib().orc_pss_taps(n_id_2, _fptr(hr), _fptr(hi))
    return hr + 1j * hi.astype(np.complex64)


def decim_taps(decim):
    t = np.zeros(4096, np.float32)
    n = lib().orc_decim_taps(decim, _fptr(t), 4096)
    return t[:n].copy()


def sss_tables(n_id_2):
    c0, c1, s, z = (np.zeros(31, np.int32) for _ in range(4))
    tab = np.zeros(900, np.int32)
    lib().orc_sss_tables(n_id_2, _iptr(c0), _iptr(c1), _iptr(s), _iptr(z), _iptr(tab))
    return c0, c1, s, z, tab.reshape(30, 30)


def cexptab():
    r, i = np.zeros(4097, np.float32), np.zeros(4097, np.float32)
    lib().orc_cexptab(_fptr(r), _fptr(i))
    return r, i


def fft128_twiddles():
    r, i = np.zeros(64, np.float32), np.zeros(64, np.float32)
    lib().orc_fft128_twiddles(_fptr(r), _fptr(i))
    return r, i


def decimate(x, decim):
    x = np.ascontiguousarray(x, np.complex64)
    n_out = (len(x) + decim - 1) // decim if decim > 1 else len(x)
    y = np.zeros(n_out, np.complex64)
    if lib().orc_decimate(x.ctypes.data, len(x), decim, y.ctypes.data) < 0:
        raise RuntimeError("orc_decimate: unsupported decimation %d" % decim)
    return y


def decimate_tcint_sc16(iq, decim=16):
    """LTB_FRONTEND_TC_INT restated: exact integer decimate-by-`decim` of [n, 2] int16 I/Q -> complex64."""
    iq = np.ascontiguousarray(iq, np.int16)
    n = iq.shape[0]
    y = np.zeros((n + decim - 1) // decim, np.complex64)
    if lib().orc_decimate_tcint_sc16_d(iq.ctypes.data, n, decim, y.ctypes.data) < 0:
        raise RuntimeError("orc_decimate_tcint_sc16 failed")
    return y


def decimate_tcint_sc8(iq, decim=16):
    """The same for [n, 2] int8 I/Q."""
    iq = np.ascontiguousarray(iq, np.int8)
    n = iq.shape[0]
    y = np.zeros((n + decim - 1) // decim, np.complex64)
    if lib().orc_decimate_tcint_sc8_d(iq.ctypes.data, n, decim, y.ctypes.data) < 0:
        raise RuntimeError("orc_decimate_tcint_sc8 failed")
    return y


def decimate_tcint_fc32(x, full_scale, decim=16):
    """The same for complex64 input taken as 23-bit fixed point over +-full_scale."""
    x = np.ascontiguousarray(x, np.complex64)
    n = x.shape[0]
    y = np.zeros((n + decim - 1) // decim, np.complex64)
    if lib().orc_decimate_tcint_fc32_d(x.ctypes.data, n, decim, float(full_scale), y.ctypes.data) < 0:
        raise RuntimeError("orc_decimate_tcint_fc32 failed")
    return y


def decimate_fast(x, decim):
    """The CPU-style evaluation used by the timed baseline (same taps, last-bit differences)."""
    x = np.ascontiguousarray(x, np.complex64)
    n_out = (len(x) + decim - 1) // decim if decim > 1 else len(x)
    y = np.zeros(n_out, np.complex64)
    if lib().orc_decimate_fast(x.ctypes.data, len(x), decim, y.ctypes.data) < 0:
        raise RuntimeError("orc_decimate_fast: unsupported decimation %d" % decim)
    return y


def sc16_to_fc32(iq, scale=1.0 / 32768.0):
    iq = np.ascontiguousarray(iq, np.int16)
    out = np.zeros(iq.size // 2, np.complex64)
    lib().orc_sc16_to_fc32(iq.ctypes.data, iq.size // 2, scale, out.ctypes.data)
    return out


def sc8_to_fc32(iq, scale=1.0 / 128.0):
    iq = np.ascontiguousarray(iq, np.int8)
    out = np.zeros(iq.size // 2, np.complex64)
    lib().orc_sc8_to_fc32(iq.ctypes.data, iq.size // 2, scale, out.ctypes.data)
    return out


def fft1024(x, inverse=False):
    """Canonical 1024-point four-step FFT of the overlap-save mode (inverse: unscaled)."""
    x = np.ascontiguousarray(x, np.complex64)
    out = np.zeros(1024, np.complex64)
    lib().orc_fft1024(x.ctypes.data, out.ctypes.data, int(inverse))
    return out


def fft1024_twiddles():
    r, i = np.zeros(1024, np.float32), np.zeros(1024, np.float32)
    lib().orc_fft1024_twiddles(_fptr(r), _fptr(i))
    return r, i


def os_filter(n_id_2):
    r, i = np.zeros(1024, np.float32), np.zeros(1024, np.float32)
    lib().orc_os_filter(n_id_2, _fptr(r), _fptr(i))
    return r, i


def pss_corr_os(x):
    """Overlap-save block powers of the whole 896-output blocks of x: [3, 896 * (len(x) // 896)]."""
    x = np.ascontiguousarray(x, np.complex64)
    n = len(x) // OS_STEP * OS_STEP
    p = np.zeros((3, max(n, 1)), np.float32)
    got = lib().orc_pss_corr_os(x.ctypes.data, len(x), p[0].ctypes.data, p[1].ctypes.data, p[2].ctypes.data)
    assert got == n
    return p[:, :n]


def pss_corr_window(win, n_id_2, conv_mode=CONV_DIRECT):
    win = np.ascontiguousarray(win, np.complex64)
    assert len(win) == HALF
    p = np.zeros(CONV_LEN, np.float32)
    lib().orc_pss_corr_window(win.ctypes.data, n_id_2, conv_mode, _fptr(p))
    return p


def pss_corr_stream(x, n_id_2):
    x = np.ascontiguousarray(x, np.complex64)
    p = np.zeros(len(x), np.float32)
    lib().orc_pss_corr_stream(x.ctypes.data, len(x), n_id_2, _fptr(p))
    return p


def fft128(x):
    x = np.ascontiguousarray(x, np.complex64)
    out = np.zeros(128, np.complex64)
    lib().orc_fft128(x.ctypes.data, out.ctypes.data)
    return out


class Pss:
    """ltetrigger.pss restated (lib/pss_impl.cc); one general_work call per work()."""

    def __init__(self, N_id_2, psr_threshold, track_after=16, track_every=8, conv_mode=CONV_DIRECT):
        self._h = lib().orc_pss_new(N_id_2, psr_threshold, track_after, track_every, conv_mode)
        if not self._h:
            raise RuntimeError("Error initializing PSS N_id_2")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_pss_free(self._h)
            self._h = None

    def work(self, buf, pos):
        """buf: complex64 array holding the stream with >=960 readable samples before `pos`
        and >=18365 after.  Returns (noutput, nconsume, out, rec)."""
        assert buf.dtype == np.complex64 and pos >= SLOT and pos + LOOKAHEAD <= len(buf)
        out = np.zeros(HALF, np.complex64)
        rec = np.zeros(1, REC_DTYPE)
        ncons = C.c_int(0)
        n = lib().orc_pss_work(self._h, buf.ctypes.data + 8 * pos, out.ctypes.data, C.byref(ncons), rec.ctypes.data)
        return n, ncons.value, out, rec[0]

    def max_psr(self): return lib().orc_pss_max_psr(self._h)
    def mean_psr(self): return lib().orc_pss_mean_psr(self._h)
    def mean_cfo(self): return lib().orc_pss_mean_cfo(self._h)
    def psr_threshold(self): return lib().orc_pss_psr_threshold(self._h)
    def set_psr_threshold(self, t): lib().orc_pss_set_psr_threshold(self._h, t)
    def tracking_score(self): return lib().orc_pss_tracking_score(self._h)


class Sss:
    """ltetrigger.sss restated (lib/sss_impl.cc)."""

    def __init__(self, N_id_2, frame_type=0):
        self._h = lib().orc_sss_new(N_id_2)
        if not self._h:
            raise RuntimeError("Error initializing SSS N_id_2")
        lib().orc_sss_set_frame_type(self._h, frame_type)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_sss_free(self._h)
            self._h = None

    def work(self, halfframe, tag_lost, rec=None):
        halfframe = np.ascontiguousarray(halfframe, np.complex64)
        out = np.zeros(HALF, np.complex64)
        r = np.zeros(1, REC_DTYPE) if rec is None else np.array([rec], REC_DTYPE)
        if rec is None:
            r["m0"] = r["m1"] = r["n_id_1"] = r["cell_id"] = -1
        lib().orc_sss_work(self._h, halfframe.ctypes.data, int(bool(tag_lost)), out.ctypes.data, r.ctypes.data)
        return out, r[0]


def chain_run(y, n_id_2, psr_threshold=4.0, track_after=16, track_every=8, conv_mode=CONV_DIRECT, stream=0):
    y = np.ascontiguousarray(y, np.complex64)
    max_recs = len(y) // (HALF - SLOT) + 2
    recs = np.zeros(max_recs, REC_DTYPE)
    n = lib().orc_chain_run(y.ctypes.data, len(y), stream, n_id_2, psr_threshold, track_after, track_every,
                            conv_mode, recs.ctypes.data, max_recs)
    if n < 0:
        raise RuntimeError("orc_chain_run failed: %d" % n)
    return recs[:n].copy()


def trigger_run(iq, decim=1, fmt=0, psr_threshold=4.0, track_after=16, track_every=8,
                conv_mode=CONV_DIRECT, nthreads=0, fc32_full_scale=0.0):
    """iq: [n_streams, n_in] complex64 (fmt 0), [n_streams, n_in, 2] int16 (fmt 1) or int8 (fmt 2).
    fc32_full_scale: with FRONT_TCINT in conv_mode, fc32 input goes through the fixed-point integer front end."""
    if fmt == 0:
        iq = np.ascontiguousarray(iq, np.complex64)
        n_streams, n_in = iq.shape
    else:
        iq = np.ascontiguousarray(iq, np.int16 if fmt == 1 else np.int8)
        n_streams, n_in = iq.shape[0], iq.shape[1]
    n_out = (n_in + decim - 1) // decim if decim > 1 else n_in
    max_recs = n_streams * 3 * (n_out // (HALF - SLOT) + 2)
    recs = np.zeros(max_recs, REC_DTYPE)
    n = lib().orc_trigger_run2(iq.ctypes.data, fmt, n_in, n_streams, decim, psr_threshold, track_after,
                               track_every, conv_mode, float(fc32_full_scale), nthreads, recs.ctypes.data, max_recs)
    if n < 0:
        raise RuntimeError("orc_trigger_run failed")
    return recs[:n].copy()
