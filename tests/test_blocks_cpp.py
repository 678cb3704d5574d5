"""The C++ block adapters (include/ltetrigger_b200_blocks.hpp) and the GNU Radio wrappers around them
(gr-ltetrigger_b200/gr_oot/lib, compiled against the stand-in GNU Radio headers of
tests/cpp/gr_stub): they compile and link against the C-ABI library on any machine; on a GPU box
the scheduler-style drivers tests/cpp/test_blocks.cpp and tests/cpp/test_gr_oot.cpp must
reproduce the oracle's restated blocks call by call."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

SRC = os.path.join(ROOT, "tests", "cpp", "test_blocks.cpp")
LIBDIR = os.path.join(ROOT, "gr-ltetrigger_b200", "lib")


OOT = os.path.join(ROOT, "gr-ltetrigger_b200", "gr_oot", "lib")
OOT_SRC = [os.path.join(ROOT, "tests", "cpp", "test_gr_oot.cpp"), os.path.join(OOT, "pss_b200_impl.cc"),
           os.path.join(OOT, "sss_b200_impl.cc")]


def build(tmp_path, gr_oot=False):
    exe = str(tmp_path / ("test_gr_oot" if gr_oot else "test_blocks"))
    inc = ["-I", os.path.join(ROOT, "include")]
    if gr_oot:
        inc += ["-I", os.path.join(ROOT, "tests", "cpp", "gr_stub"), "-I", OOT]
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall"] + inc + (OOT_SRC if gr_oot else [SRC]) +
                          ["-L", LIBDIR, "-lltetrigger_b200", "-Wl,-rpath," + LIBDIR, "-o", exe])
    return exe


@pytest.mark.parametrize("gr_oot", [False, True])
def test_adapters_compile_and_link(tmp_path, gr_oot):
    exe = build(tmp_path, gr_oot)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 2 and "usage" in out.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("gr_oot", [False, True, "hier"])
def test_cpp_blocks_match_oracle(tmp_path, oracle, gr_oot):
    """gr_oot = "hier": the GNU Radio wrappers of all three chains on ONE shared engine (LTB_SHARE_ENGINE=1),
    driven under GNU Radio's input-buffer limit; chain 0's calls must still equal the oracle's."""
    exe = build(tmp_path, bool(gr_oot))
    fixture = os.path.join(GOLDEN, "test_frames", "lte_frame_6prb_cellid_123")
    env = dict(os.environ, LTB_SHARE_ENGINE="1") if gr_oot == "hier" else dict(os.environ)
    args = [exe, fixture, "0.4", "0", "4"] + (["hier"] if gr_oot == "hier" else [])
    out = subprocess.run(args, capture_output=True, text=True, check=True, env=env).stdout.splitlines()
    assert out[0] == "E Error initializing PSS N_id_2"                      # lib/pss_impl.cc:75-76
    x = np.fromfile(fixture, np.complex64)
    n = int(0.4 * 1.92e6) // 8 * 8
    x = np.tile(x, -(-n // len(x)))[:n]
    op, os_ = oracle.Pss(0, 4.0), oracle.Sss(0)
    buf = np.concatenate([np.zeros(960, np.complex64), x])
    pos, written = 960, 0
    lines = iter(out[1:])
    n_calls = 0
    while pos - 960 + oracle.LOOKAHEAD <= n:
        nout, ncons, o, rec = op.work(buf, pos)
        lost = int(bool(rec["flags"] & oracle.F_TAG_LOST))
        assert next(lines) == "P %d %d %d %d" % (pos - 960, nout, ncons, lost)
        if nout:
            _, srec = os_.work(o, lost)
            cell = int(srec["cell_id"]) if srec["flags"] & oracle.F_CELL else -1
            cp = int(bool(srec["flags"] & oracle.F_CP_NORM)) if srec["flags"] & oracle.F_CELL else -1
            s = 0
            for w in o.view(np.uint32).tolist():
                s = (s * 1000003 + w) & 0xFFFFFFFFFFFFFFFF
            assert next(lines) == "S %d %d %d %d" % (written, cell, cp, s)
            written += nout
        pos += ncons
        n_calls += 1
    assert n_calls >= 70
    acc = next(lines).split()
    assert acc[0] == "A"
    want = [op.max_psr(), op.mean_psr(), op.mean_cfo(), op.psr_threshold(), op.tracking_score()]
    assert [np.float32(v) for v in acc[1:]] == [np.float32(v) for v in want]
