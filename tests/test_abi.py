"""CPU-side checks of the C ABI: the library loads, exports every symbol the header
declares, its host-side tables equal the oracle's bit for bit, and compute entry points
fail loudly (never fall back) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_gpu


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ltetrigger_b200.h")).read()
    text = re.sub(r"#ifdef LTB_DEBUG.*?#endif", "", text, flags=re.S)        # debug-build-only declarations
    return sorted(set(re.findall(r"LTB_API\s+[\w\s\*]+?\b(ltb_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import ltetrigger_b200 as lt
    from ltetrigger_b200 import _abi
    L = lt.lib()
    declared = header_symbols()
    assert len(declared) >= 20
    assert sorted(_abi.SYMBOLS) == declared
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.ltb_version()
    # the process-wide debug switch exists only in the -DLTB_DEBUG build, never in the release library
    assert not hasattr(L, "ltb_debug_set_flag")
    assert hasattr(_abi.debug_lib(), "ltb_debug_set_flag")


def test_struct_layouts():
    from ltetrigger_b200 import _abi
    assert _abi.WINDOW_REC.itemsize == 88
    assert C.sizeof(_abi.TriggerConfig) == 88 and _abi.TriggerConfig.corr_mode.offset == 64
    assert _abi.TriggerConfig.frontend_mode.offset == 72 and _abi.TriggerConfig.pipeline.offset == 76
    assert _abi.TriggerConfig.fc32_full_scale.offset == 80
    assert C.sizeof(_abi.PssStats) == 32
    from oracle import oracle as O
    assert O.REC_DTYPE == _abi.WINDOW_REC


def test_tables_match_oracle(oracle):
    import ltetrigger_b200 as lt
    for r in range(3):
        assert np.array_equal(lt.tables.pss_taps(r).view(np.uint32), oracle.pss_taps(r).view(np.uint32))
        for a, b in zip(lt.tables.sss(r), oracle.sss_tables(r)):
            assert np.array_equal(a, b)
    for d in range(2, lt.MAX_DECIM + 1):
        assert np.array_equal(lt.tables.decim_taps(d).view(np.uint32), oracle.decim_taps(d).view(np.uint32))
    for a, b in zip(lt.tables.cexp(), oracle.cexptab()):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    for a, b in zip(lt.tables.fft128_twiddles(), oracle.fft128_twiddles()):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    for a, b in zip(lt.tables.fft1024_twiddles(), oracle.fft1024_twiddles()):      # overlap-save mode
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    for r in range(3):
        for a, b in zip(lt.tables.os_filter(r), oracle.os_filter(r)):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_invalid_inputs():
    import ltetrigger_b200 as lt
    from ltetrigger_b200 import _abi
    L = lt.lib()
    h = C.c_void_p()
    cfg = _abi.TriggerConfig()
    assert L.ltb_trigger_create(C.byref(cfg), C.byref(h)) == lt.ERROR_INVALID_INPUTS    # struct_size 0
    cfg.struct_size = C.sizeof(_abi.TriggerConfig)
    cfg.n_streams, cfg.decim, cfg.max_chunk = 1, lt.MAX_DECIM + 1, 9600
    assert L.ltb_trigger_create(C.byref(cfg), C.byref(h)) == lt.ERROR_INVALID_INPUTS    # decim 65
    cfg.decim, cfg.input_format = 3, 7
    assert L.ltb_trigger_create(C.byref(cfg), C.byref(h)) == lt.ERROR_INVALID_INPUTS    # unknown format
    # ABI evolution: the config as it was before corr_mode was appended (64 bytes) is still accepted
    # (validation passes; without a GPU the call then fails on the device, with one it succeeds)
    cfg.decim, cfg.input_format, cfg.struct_size = 1, 0, 64
    rc = L.ltb_trigger_create(C.byref(cfg), C.byref(h))
    assert rc != lt.ERROR_INVALID_INPUTS
    if rc == lt.SUCCESS:
        L.ltb_trigger_destroy(h)
    cfg.struct_size = 60
    assert L.ltb_trigger_create(C.byref(cfg), C.byref(h)) == lt.ERROR_INVALID_INPUTS
    cfg.struct_size, cfg.corr_mode = C.sizeof(_abi.TriggerConfig), 5
    assert L.ltb_trigger_create(C.byref(cfg), C.byref(h)) == lt.ERROR_INVALID_INPUTS    # unknown correlator
    re_, im_ = np.zeros(128, np.float32), np.zeros(128, np.float32)
    assert L.ltb_table_pss_taps(3, _abi.fptr(re_), _abi.fptr(im_)) == lt.ERROR_INVALID_INPUTS
    s = C.c_void_p()
    assert L.ltb_sss_create(0, 5, C.byref(s)) == lt.ERROR_INVALID_INPUTS


@pytest.mark.skipif(has_gpu(), reason="checks the no-device behaviour")
def test_no_device_fails_loudly():
    import ltetrigger_b200 as lt
    assert lt.device_count() == 0
    with pytest.raises(lt.LtbError):
        lt.Trigger(n_streams=1)
    with pytest.raises(lt.LtbError):
        lt.kernel_pss_corr(np.zeros((1, 64), np.complex64))
    with pytest.raises(RuntimeError):
        lt.sss(0)


def _build_c_client(tmp_path):
    import subprocess
    libdir = os.path.join(ROOT, "gr-ltetrigger_b200", "lib")
    exe = str(tmp_path / "test_abi_c")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "test_abi.c"), "-L", libdir, "-lltetrigger_b200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    return exe


@pytest.mark.skipif(has_gpu(), reason="checks the no-device behaviour")
def test_c_client_compiles_links_and_fails_loudly_without_a_device(tmp_path):
    """include/ltetrigger_b200.h is valid C99 and the library is usable from plain C."""
    import subprocess
    out = subprocess.run([_build_c_client(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "taps 525" in out.stdout and "no device: create -> -1" in out.stdout


@pytest.mark.gpu
def test_c_client_finds_the_fixture_cell(tmp_path):
    import subprocess
    fixture = os.path.join(ROOT, "tests", "golden", "test_frames", "lte_frame_6prb_cellid_123")
    out = subprocess.run([_build_c_client(tmp_path), fixture], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "cell_id 123" in out.stdout


def test_built_library_holds_the_tensor_core_and_tma_paths():
    """The shipped .so, disassembled (cuobjdump -sass; tools/sass_evidence.py): every variant of the integer front end
    issues tcgen05.mma kind::i8 (UTCIMMA), reads its accumulators from tensor memory (LDTM), loads its operands by TMA
    (UTMALDG) and signals completion through tcgen05.commit (UTCBAR) and mbarriers (SYNCS); the FP32 streaming decimators
    stage their input with bulk copies (UBLKCP) and compute in packed FFMA2.  No mma.sync / wgmma-era HMMA anywhere."""
    import shutil
    import subprocess
    import sys
    if not shutil.which("cuobjdump") or not shutil.which("cu++filt"):
        pytest.skip("CUDA binary utilities not installed")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_evidence
    k = sass_evidence.per_kernel()
    tc = {n: c for n, c in k.items() if "decimate_tc_kernel<" in n}
    assert len(tc) == 17, sorted(tc)                 # fc32 at 7 rates (D = 2 ... 32), sc16 at 6 (from D = 4), sc8 at 4 (from D = 8)
    for n, c in tc.items():
        assert c["UTCIMMA"] >= 4 and c["LDTM"] >= 1 and c["UTMALDG"] >= 1 and c["UTCBAR"] >= 1 and c["SYNCS"] >= 8, (n, dict(c))
        assert c["FFMA2"] == 0, n
    for fmt in range(3):
        c = k["ltb::decimate_stream_kernel<%d>" % fmt]
        assert c["UBLKCP"] >= 1 and c["SYNCS"] >= 1 and c["FFMA2"] >= 500, dict(c)
    assert k["ltb::pss_corr_fft_kernel"]["FFMA2"] > 100 and k["ltb::pss_track_kernel"]["FFMA2"] > 100
    sass = subprocess.run(["cuobjdump", "-sass", sass_evidence.LIB], capture_output=True, text=True).stdout
    assert " HMMA" not in sass and "WGMMA" not in sass
