"""ctypes binding of libltetrigger_b200.so (include/ltetrigger_b200.h).

The library is the product; this module only marshals pointers.  There is no CPU
fallback: if the shared library is missing, import fails; if no CUDA device is present,
every compute call returns LTB_ERROR and the wrappers raise.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(_PKG, "..", ".."))          # gr-ltetrigger_b200/
# LTB200_LIB overrides the in-tree build (e.g. an instrumented build while profiling)
LIB_PATH = os.environ.get("LTB200_LIB") or os.path.join(ROOT, "lib", "libltetrigger_b200.so")

SUCCESS, ERROR, ERROR_INVALID_INPUTS = 0, -1, -2
SLOT_LEN, HALF_FRAME, SYMBOL_SZ, CONV_LEN, LOOKAHEAD = 960, 9600, 128, 9726, 18365
FMT_FC32, FMT_SC16, FMT_SC8 = 0, 1, 2
MAX_DECIM = 64
CORR_DIRECT, CORR_FFT, OS_STEP = 0, 1, 896
FRAME_FDD, FRAME_TDD = 0, 1
FRONTEND_FP32, FRONTEND_TC_INT = 0, 1
PIPE_OVERLAP, PIPE_SERIAL = 0, 1
FMT_BYTES = {FMT_FC32: 8, FMT_SC16: 4, FMT_SC8: 2}
FMT_DTYPE = {FMT_FC32: np.complex64, FMT_SC16: np.int16, FMT_SC8: np.int8}
MIN_PSR_THRESHOLD = 1.5
F_SEARCHED, F_OVER, F_EMIT, F_TRACKING, F_TAG_LOST, F_SSS, F_CELL, F_CP_NORM = (
    0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80)

WINDOW_REC = np.dtype([
    ("win_start", "<i8"), ("emit_start", "<i8"), ("stream", "<i4"), ("n_id_2", "<i4"),
    ("win_index", "<i4"), ("flags", "<u4"), ("peak_pos", "<i4"), ("score", "<i4"),
    ("psr", "<f4"), ("peak_value", "<f4"), ("cfo", "<f4"), ("mean_cfo", "<f4"),
    ("m0", "<i4"), ("m1", "<i4"), ("m0_val", "<f4"), ("m1_val", "<f4"),
    ("n_id_1", "<i4"), ("cell_id", "<i4"), ("cp_norm_avg", "<f4"), ("cp_ext_avg", "<f4")],
    align=True)
assert WINDOW_REC.itemsize == 88


class TriggerConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("n_streams", C.c_int32),
                ("input_format", C.c_int32), ("decim", C.c_int32), ("root_mask", C.c_int32),
                ("max_chunk", C.c_int64), ("psr_threshold", C.c_float), ("track_after", C.c_int32),
                ("track_every", C.c_int32), ("record_all", C.c_int32), ("keep_halfframes", C.c_int32),
                ("cuda_stream", C.c_void_p), ("corr_mode", C.c_int32), ("frame_type", C.c_int32),
                ("frontend_mode", C.c_int32), ("pipeline", C.c_int32), ("fc32_full_scale", C.c_float)]


class Mib(C.Structure):
    _fields_ = [("nof_prb", C.c_int32), ("nof_ports", C.c_int32), ("phich_length", C.c_int32),
                ("phich_resources", C.c_int32), ("sfn", C.c_int32), ("sfn_offset", C.c_int32)]


class PssStats(C.Structure):
    _fields_ = [("max_psr", C.c_float), ("mean_psr", C.c_float), ("mean_cfo", C.c_float),
                ("psr_threshold", C.c_float), ("tracking_score", C.c_float), ("tracking", C.c_int32),
                ("next_window", C.c_int64)]


# every symbol include/ltetrigger_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "ltb_trigger_create", "ltb_trigger_destroy", "ltb_trigger_reset", "ltb_trigger_set_psr_threshold",
    "ltb_trigger_process_host", "ltb_trigger_process_device", "ltb_trigger_submit_device", "ltb_trigger_submit_host",
    "ltb_trigger_collect", "ltb_trigger_get_stats", "ltb_trigger_fetch_halfframes",
    "ltb_trigger_last_timing", "ltb_trigger_last_kernel_times", "ltb_last_error", "ltb_version", "ltb_device_count",
    "ltb_sss_create", "ltb_sss_destroy", "ltb_sss_set_frame_type", "ltb_sss_work", "ltb_mib_decode",
    "ltb_kernel_pss_corr_host", "ltb_kernel_pss_corr_fft_host", "ltb_kernel_decimate_host", "ltb_kernel_decimate_tc_host",
    "ltb_table_pss_taps", "ltb_table_decim_taps", "ltb_table_sss", "ltb_table_cexp",
    "ltb_table_fft128_twiddles", "ltb_table_fft1024_twiddles", "ltb_table_os_filter", "ltb_table_tc_btab",
]

DEBUG_LIB_PATH = os.path.join(ROOT, "lib", "libltetrigger_b200_debug.so")

_lib = None
_debug_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load(LIB_PATH)
    return _lib


def debug_lib():
    """The -DLTB_DEBUG build of the same sources: exports ltb_debug_set_flag (decimator dissection and
    kernel selection).  Only tests and tools/dissect.py load it; the release library has no such switch."""
    global _debug_lib
    if _debug_lib is None:
        _debug_lib = _load(DEBUG_LIB_PATH)
        _debug_lib.ltb_debug_set_flag.argtypes = [C.c_int, C.c_int]
    return _debug_lib


def _load(path):
    if not os.path.exists(path):
        raise ImportError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(or make -C gr-ltetrigger_b200)" % path)
    L = C.CDLL(path)
    vp, fp, ip = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)
    L.ltb_last_error.restype = C.c_char_p
    L.ltb_version.restype = C.c_char_p
    L.ltb_trigger_create.argtypes = [C.POINTER(TriggerConfig), C.POINTER(vp)]
    L.ltb_trigger_destroy.argtypes = [vp]
    L.ltb_trigger_reset.argtypes = [vp]
    L.ltb_trigger_set_psr_threshold.argtypes = [vp, C.c_int, C.c_int, C.c_float, C.c_int]
    L.ltb_trigger_process_host.argtypes = [vp, vp, C.c_int64, C.c_int64, vp, C.c_int, ip]
    L.ltb_trigger_process_device.argtypes = [vp, vp, C.c_int64, C.c_int64, vp, C.c_int, ip]
    L.ltb_trigger_submit_device.argtypes = [vp, vp, C.c_int64, C.c_int64]
    L.ltb_trigger_submit_host.argtypes = [vp, vp, C.c_int64, C.c_int64]
    L.ltb_trigger_collect.argtypes = [vp, vp, C.c_int, ip]
    L.ltb_trigger_get_stats.argtypes = [vp, C.c_int, C.c_int, C.POINTER(PssStats)]
    L.ltb_trigger_fetch_halfframes.argtypes = [vp, vp, C.c_int, ip]
    L.ltb_trigger_last_timing.argtypes = [vp, fp, ip]
    L.ltb_trigger_last_kernel_times.argtypes = [vp, fp]
    L.ltb_sss_create.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
    L.ltb_sss_destroy.argtypes = [vp]
    L.ltb_sss_set_frame_type.argtypes = [vp, C.c_int]
    L.ltb_sss_work.argtypes = [vp, vp, vp, C.c_int, vp]
    L.ltb_mib_decode.argtypes = [vp, C.c_int, C.c_int, C.POINTER(Mib)]
    L.ltb_kernel_pss_corr_host.argtypes = [C.c_int, vp, C.c_int, C.c_int64, vp]
    L.ltb_kernel_pss_corr_fft_host.argtypes = [C.c_int, vp, C.c_int, C.c_int64, vp]
    L.ltb_table_fft1024_twiddles.argtypes = [fp, fp]
    L.ltb_table_os_filter.argtypes = [C.c_int, fp, fp]
    L.ltb_table_tc_btab.argtypes = [C.c_int, C.c_int, vp, vp]
    L.ltb_kernel_decimate_host.argtypes = [C.c_int, vp, C.c_int, C.c_int, C.c_int64, C.c_int, vp]
    L.ltb_kernel_decimate_tc_host.argtypes = [C.c_int, vp, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int64, C.c_int64, vp]
    L.ltb_table_pss_taps.argtypes = [C.c_int, fp, fp]
    L.ltb_table_decim_taps.argtypes = [C.c_int, fp, C.c_int]
    L.ltb_table_sss.argtypes = [C.c_int, ip, ip, ip, ip, ip]
    L.ltb_table_cexp.argtypes = [fp, fp]
    L.ltb_table_fft128_twiddles.argtypes = [fp, fp]
    return L


class LtbError(RuntimeError):
    pass


def check(rc, what):
    if rc != SUCCESS:
        msg = lib().ltb_last_error().decode(errors="replace")
        raise LtbError("%s failed (%d): %s" % (what, rc, msg))


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))
