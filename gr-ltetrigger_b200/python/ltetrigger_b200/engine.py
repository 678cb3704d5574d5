"""Batched trigger engine: thin object wrapper over the ltb_trigger_* C ABI."""
import ctypes as C

import numpy as np

from . import _abi as A


def device_count():
    return A.lib().ltb_device_count()


class Trigger:
    """n_streams x 3 chains of pss(N_id_2=k) -> sss(N_id_2=k), fed in chunks.

    Stands where `rational_resampler_ccc(1, decim) -> downlink_trigger_c(psr_threshold)`
    stands in the reference's flowgraphs (examples/cell_search_file.py:56-60), minus mib.
    """

    def __init__(self, n_streams, decim=1, psr_threshold=4.0, max_chunk=1 << 20, input_format=A.FMT_FC32,
                 track_after=16, track_every=8, record_all=True, keep_halfframes=False, device=0,
                 root_mask=7, cuda_stream=None, corr_mode=A.CORR_DIRECT, frame_type=A.FRAME_FDD,
                 frontend_mode=A.FRONTEND_FP32, pipeline=A.PIPE_OVERLAP, fc32_full_scale=0.0):
        cfg = A.TriggerConfig()
        cfg.struct_size = C.sizeof(A.TriggerConfig)
        cfg.device, cfg.n_streams, cfg.input_format, cfg.decim = device, n_streams, input_format, decim
        cfg.root_mask, cfg.max_chunk, cfg.psr_threshold = root_mask, max_chunk, psr_threshold
        cfg.track_after, cfg.track_every = track_after, track_every
        cfg.record_all, cfg.keep_halfframes = int(record_all), int(keep_halfframes)
        cfg.cuda_stream = cuda_stream
        cfg.corr_mode = corr_mode
        cfg.frame_type = frame_type
        cfg.frontend_mode = frontend_mode
        cfg.pipeline = pipeline
        cfg.fc32_full_scale = fc32_full_scale
        self._h = C.c_void_p()
        A.check(A.lib().ltb_trigger_create(C.byref(cfg), C.byref(self._h)), "ltb_trigger_create")
        self.n_streams, self.decim, self.input_format = n_streams, decim, input_format
        self.max_chunk = max_chunk
        self._recs = np.zeros(n_streams * 3 * (max_chunk // decim // 8640 + 6), A.WINDOW_REC)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            A.lib().ltb_trigger_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self):
        A.check(A.lib().ltb_trigger_reset(self._h), "ltb_trigger_reset")

    def set_psr_threshold(self, thr, stream=-1, n_id_2=-1, clamp=True):
        A.check(A.lib().ltb_trigger_set_psr_threshold(self._h, stream, n_id_2, thr, int(clamp)),
                "ltb_trigger_set_psr_threshold")

    def _bytes_per_sample(self):
        return A.FMT_BYTES[self.input_format]

    def process(self, iq):
        """iq: host array [n_streams, n] complex64 (fc32), [n_streams, n, 2] int16 (sc16) or int8
        (sc8).  Returns the window records of this chunk as a WINDOW_REC array."""
        iq = np.ascontiguousarray(iq, A.FMT_DTYPE[self.input_format])
        assert iq.shape[0] == self.n_streams
        n = iq.shape[1]
        nrec = C.c_int32(0)
        rc = A.lib().ltb_trigger_process_host(self._h, iq.ctypes.data, n * self._bytes_per_sample(), n,
                                              self._recs.ctypes.data, len(self._recs), C.byref(nrec))
        A.check(rc, "ltb_trigger_process_host")
        return self._recs[:nrec.value].copy()

    def process_host_ptr(self, ptr, stride_bytes, n):
        nrec = C.c_int32(0)
        rc = A.lib().ltb_trigger_process_host(self._h, ptr, stride_bytes, n, self._recs.ctypes.data,
                                              len(self._recs), C.byref(nrec))
        A.check(rc, "ltb_trigger_process_host")
        return self._recs[:nrec.value]

    def process_device_ptr(self, dptr, stride_bytes, n):
        nrec = C.c_int32(0)
        rc = A.lib().ltb_trigger_process_device(self._h, dptr, stride_bytes, n, self._recs.ctypes.data,
                                                len(self._recs), C.byref(nrec))
        A.check(rc, "ltb_trigger_process_device")
        return self._recs[:nrec.value]

    def submit_device_ptr(self, dptr, stride_bytes, n):
        A.check(A.lib().ltb_trigger_submit_device(self._h, dptr, stride_bytes, n), "ltb_trigger_submit_device")

    def submit_host_ptr(self, ptr, stride_bytes, n):
        A.check(A.lib().ltb_trigger_submit_host(self._h, ptr, stride_bytes, n), "ltb_trigger_submit_host")

    def collect(self):
        nrec = C.c_int32(0)
        A.check(A.lib().ltb_trigger_collect(self._h, self._recs.ctypes.data, len(self._recs), C.byref(nrec)),
                "ltb_trigger_collect")
        return self._recs[:nrec.value]

    def run(self, iq, chunk=None):
        """Feed a whole capture in chunks; returns all records ordered (stream, n_id_2, win_index)."""
        n = iq.shape[1]
        step = 8 * self.decim
        chunk = min(chunk or self.max_chunk, self.max_chunk) // step * step
        out = []
        for a in range(0, n - n % step, chunk):
            out.append(self.process(iq[:, a:min(a + chunk, n - n % step)]))
        recs = np.concatenate(out) if out else np.zeros(0, A.WINDOW_REC)
        order = np.lexsort((recs["win_index"], recs["n_id_2"], recs["stream"]))
        return recs[order]

    def stats(self, stream, n_id_2):
        st = A.PssStats()
        A.check(A.lib().ltb_trigger_get_stats(self._h, stream, n_id_2, C.byref(st)), "ltb_trigger_get_stats")
        return st

    def fetch_halfframes(self, max_halfframes):
        out = np.zeros((max_halfframes, A.HALF_FRAME), np.complex64)
        n = C.c_int32(0)
        A.check(A.lib().ltb_trigger_fetch_halfframes(self._h, out.ctypes.data, max_halfframes, C.byref(n)),
                "ltb_trigger_fetch_halfframes")
        return out[:n.value]

    def last_timing(self):
        ms, nl = C.c_float(0), C.c_int32(0)
        A.lib().ltb_trigger_last_timing(self._h, C.byref(ms), C.byref(nl))
        return ms.value, nl.value

    def last_kernel_times(self):
        """ms of the last call's [front end, PSS correlator, track, SSS] stages (CUDA events)."""
        ms = (C.c_float * 4)()
        A.lib().ltb_trigger_last_kernel_times(self._h, ms)
        return list(ms)


def kernel_pss_corr(x, device=0):
    """x: [n_streams, n] complex64 -> power [n_streams, 3, n] (sliding, x[<0]=0)."""
    x = np.ascontiguousarray(np.atleast_2d(x), np.complex64)
    s, n = x.shape
    p = np.zeros((s, 3, n), np.float32)
    A.check(A.lib().ltb_kernel_pss_corr_host(device, x.ctypes.data, s, n, p.ctypes.data), "ltb_kernel_pss_corr_host")
    return p


def kernel_pss_corr_fft(x, device=0):
    """Overlap-save FFT evaluation: x [n_streams, n] complex64 -> power [n_streams, 3, 896 * (n // 896)]."""
    x = np.ascontiguousarray(np.atleast_2d(x), np.complex64)
    s, n = x.shape
    p = np.zeros((s, 3, n // A.OS_STEP * A.OS_STEP), np.float32)
    A.check(A.lib().ltb_kernel_pss_corr_fft_host(device, x.ctypes.data, s, n, p.ctypes.data), "ltb_kernel_pss_corr_fft_host")
    return p


def kernel_decimate(x, decim, fmt=A.FMT_FC32, device=0, L=None):
    """L: the library to call (default the release build; tests pass A.debug_lib())."""
    if fmt == A.FMT_FC32:
        x = np.ascontiguousarray(np.atleast_2d(x), np.complex64)
    else:
        x = np.ascontiguousarray(x, A.FMT_DTYPE[fmt])
    s, n = x.shape[0], x.shape[1]
    y = np.zeros((s, n // decim), np.complex64)
    A.check((L or A.lib()).ltb_kernel_decimate_host(device, x.ctypes.data, fmt, s, n, decim, y.ctypes.data),
            "ltb_kernel_decimate_host")
    return y


def kernel_decimate_tc(iq, chunk=None, device=0, full_scale=0.0, decim=16):
    """LTB_FRONTEND_TC_INT at kernel level: iq [n_streams, n, 2] int16 (sc16) or int8 (sc8), or [n_streams, n]
    complex64 taken as 23-bit fixed point over +-full_scale -> [n_streams, n // decim] complex64, the input fed in
    calls of `chunk` samples (multiple of 8 decim; default: one call)."""
    iq = np.asarray(iq)
    fmt = A.FMT_SC8 if iq.dtype == np.int8 else A.FMT_SC16 if iq.dtype == np.int16 else A.FMT_FC32
    iq = np.ascontiguousarray(iq, A.FMT_DTYPE[fmt])
    s, n = iq.shape[0], iq.shape[1]
    y = np.zeros((s, n // decim), np.complex64)
    A.check(A.lib().ltb_kernel_decimate_tc_host(device, iq.ctypes.data, fmt, decim, full_scale, s, n, chunk or n, y.ctypes.data),
            "ltb_kernel_decimate_tc_host")
    return y


class tables:
    """Host-side constant tables (no GPU needed)."""

    @staticmethod
    def pss_taps(n_id_2):
        re, im = np.zeros(128, np.float32), np.zeros(128, np.float32)
        A.check(A.lib().ltb_table_pss_taps(n_id_2, A.fptr(re), A.fptr(im)), "ltb_table_pss_taps")
        return re + 1j * im.astype(np.complex64)

    @staticmethod
    def decim_taps(decim):
        t = np.zeros(4096, np.float32)
        n = A.lib().ltb_table_decim_taps(decim, A.fptr(t), 4096)
        if n < 0:
            raise A.LtbError("ltb_table_decim_taps failed")
        return t[:n].copy()

    @staticmethod
    def sss(n_id_2):
        c0, c1, s, z = (np.zeros(31, np.int32) for _ in range(4))
        tab = np.zeros(900, np.int32)
        A.check(A.lib().ltb_table_sss(n_id_2, A.iptr(c0), A.iptr(c1), A.iptr(s), A.iptr(z), A.iptr(tab)), "ltb_table_sss")
        return c0, c1, s, z, tab.reshape(30, 30)

    @staticmethod
    def cexp():
        r, i = np.zeros(4097, np.float32), np.zeros(4097, np.float32)
        A.check(A.lib().ltb_table_cexp(A.fptr(r), A.fptr(i)), "ltb_table_cexp")
        return r, i

    @staticmethod
    def fft1024_twiddles():
        r, i = np.zeros(1024, np.float32), np.zeros(1024, np.float32)
        A.check(A.lib().ltb_table_fft1024_twiddles(A.fptr(r), A.fptr(i)), "ltb_table_fft1024_twiddles")
        return r, i

    @staticmethod
    def os_filter(n_id_2):
        r, i = np.zeros(1024, np.float32), np.zeros(1024, np.float32)
        A.check(A.lib().ltb_table_os_filter(n_id_2, A.fptr(r), A.fptr(i)), "ltb_table_os_filter")
        return r, i

    @staticmethod
    def tc_btab(fmt, decim=16):
        """Tap table of the tensor-core front end, un-swizzled: ([208, 128] int8, sum of the integer taps)."""
        raw = np.zeros(208 * 128, np.int8)
        sum_t = C.c_int64(0)
        A.check(A.lib().ltb_table_tc_btab(fmt, decim, raw.ctypes.data, C.addressof(sum_t)), "ltb_table_tc_btab")
        raw = raw.reshape(208, 8, 16)
        out = np.zeros_like(raw)
        for r in range(208):
            for c in range(8):
                out[r, c] = raw[r, c ^ (r & 7)]
        return out.reshape(208, 128), int(sum_t.value)

    @staticmethod
    def fft128_twiddles():
        r, i = np.zeros(64, np.float32), np.zeros(64, np.float32)
        A.check(A.lib().ltb_table_fft128_twiddles(A.fptr(r), A.fptr(i)), "ltb_table_fft128_twiddles")
        return r, i
