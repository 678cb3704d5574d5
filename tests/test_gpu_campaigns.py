"""The long parity campaigns, inside `pytest -m gpu` so that the driver's round-end run sees them:
BASELINE config C4 at full size (256 streams x 21 SNR points, both matched-filter evaluations), a 60 s
seeded fuzz campaign over engine configurations, and one C5-shard-sized call (512 streams x 100 ms at
30.72 Msps, decimate by 16) -- every window record compared bit for bit with the CPU oracle."""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np
import pytest

from conftest import ROOT, assert_recs_equal

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(ROOT, "examples"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


@pytest.fixture(scope="module")
def lt():
    import ltetrigger_b200 as lt
    if lt.device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU path")
    return lt


def _save(name, obj):
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, name), "w") as f:
            json.dump(obj, f, indent=1)


def test_c4_snr_sweep_full_size_both_correlators(lt, oracle):
    """Config C4 (BASELINE.json configs[3]; the batch form of examples/snr_ltetrigger.grc): 256 seeded
    streams x 0.5 s at 1.92 Msps per SNR point, -10..+10 dB in 1 dB steps, cell ids dealt from a
    permutation of 0..503, threshold 4.  Both correlators against the oracle's restatement of the
    same arithmetic; detection counts identical in both modes; no stream tags a wrong cell."""
    import snr_sweep
    S, n, thr, seed = 256, 960000, 4.0, 20260
    pool = mp.get_context("fork").Pool(min(os.cpu_count() or 1, 32))
    trigs = {"fft": lt.Trigger(n_streams=S, decim=1, psr_threshold=thr, max_chunk=n, corr_mode=lt.CORR_FFT),
             "direct": lt.Trigger(n_streams=S, decim=1, psr_threshold=thr, max_chunk=n, corr_mode=lt.CORR_DIRECT)}
    conv = {"fft": oracle.CONV_OS, "direct": oracle.CONV_DIRECT}
    points = []
    try:
        for snr in range(-10, 11):
            iq, ids = snr_sweep.make_batch(pool, S, n, float(snr), seed + 1000 * (snr + 100))
            pt = {"snr_db": snr, "streams": S}
            for name, trig in trigs.items():
                trig.reset()
                got = trig.run(iq)
                want = oracle.trigger_run(iq, decim=1, psr_threshold=thr, conv_mode=conv[name])
                assert_recs_equal(got, want)
                tagged = got[(got["flags"] & lt.F_CELL) != 0]
                right = wrong = 0
                for s in range(S):
                    c = tagged[tagged["stream"] == s]["cell_id"]
                    if len(c):
                        if np.bincount(c).argmax() == ids[s]:
                            right += 1
                        else:
                            wrong += 1
                pt[name] = {"detected": right, "wrong_cell": wrong, "records": int(len(got))}
            assert pt["fft"]["detected"] == pt["direct"]["detected"], pt
            assert pt["fft"]["wrong_cell"] == 0 and pt["direct"]["wrong_cell"] == 0, pt
            points.append(pt)
    finally:
        pool.close()
        for t in trigs.values():
            t.close()
    det = {p["snr_db"]: p["fft"]["detected"] for p in points}
    assert det[-10] == 0 and det[10] >= 250 and det[0] >= 240, det
    _save("c4_snr_sweep_pytest.json", {"config": "C4: 256 streams x 0.5 s x 21 SNR points, threshold 4, both correlators, "
                                       "every record bit-identical to the oracle", "points": points})


def test_fuzz_parity_60s(lt):
    """tests/fuzz_parity.py for 60 s with a fixed seed: random rate / format / correlator / frame type /
    chunking / thresholds / tracking parameters / SNR / CP / CFO, every record against the oracle."""
    import fuzz_parity
    lines = []
    n_cases, bad = fuzz_parity.campaign(60.0, seed=3, log=lines.append)
    _save("fuzz_parity_pytest.json", {"seed": 3, "seconds": 60, "cases": n_cases, "mismatch": bad, "log_tail": lines[-5:]})
    assert bad is None, bad
    assert n_cases >= 20, n_cases


@pytest.mark.parametrize("frontend", ["fp32", "tc"])
def test_c5_shard_one_call_512_streams_vs_oracle(lt, oracle, frontend):
    """One bench-sized call -- 512 streams x 100 ms x 30.72 Msps fc32, decimate by 16, FFT correlator,
    fed from device memory like bench.py, with the canonical FP32 front end and with the integer tensor-core
    one (fc32 as fixed point over +-8, bench.py's setting) -- with ALL records compared with the oracle in the
    same mode (run on the host in four groups of 128 streams to bound host memory)."""
    import torch
    S, decim, n, U = 512, 16, 16 * 192000, 8
    from ltetrigger_b200 import synth
    dev = torch.device("cuda", 0)
    perm = np.random.default_rng(2027).permutation(504)
    base = torch.from_numpy(np.stack([synth.capture(int(perm[u]), n, snr_db=None, decim=decim, seed=91 + u, offset=0)
                                      for u in range(U)])).to(dev)
    g = torch.Generator(device=dev)
    g.manual_seed(4242)
    shifts = torch.randint(0, n, (S,), generator=torch.Generator().manual_seed(9))
    snr = np.linspace(-6.0, 12.0, S)
    x = torch.empty((S, n), dtype=torch.complex64, device=dev)
    for s in range(S):
        sigma = float(np.sqrt(1.0 / (10.0 ** (snr[s] / 10.0)) / 2.0))
        noise = torch.randn((n, 2), generator=g, device=dev, dtype=torch.float32)
        x[s] = torch.roll(base[s % U], int(shifts[s])) + sigma * torch.view_as_complex(noise)
    del base, noise
    torch.cuda.synchronize()
    tc = frontend == "tc"
    trig = lt.Trigger(n_streams=S, decim=decim, psr_threshold=4.0, max_chunk=n, device=0, corr_mode=lt.CORR_FFT,
                      frontend_mode=lt.FRONTEND_TC_INT if tc else lt.FRONTEND_FP32, fc32_full_scale=8.0 if tc else 0.0)
    got = trig.process_device_ptr(x.data_ptr(), n * 8, n).copy()
    trig.close()
    got = got[np.lexsort((got["win_index"], got["n_id_2"], got["stream"]))]
    assert len(got) > 15 * S
    t0 = time.time()
    for g0 in range(0, S, 128):
        iq = x[g0:g0 + 128].cpu().numpy()
        want = oracle.trigger_run(iq, decim=decim, psr_threshold=4.0, conv_mode=oracle.CONV_OS | (oracle.FRONT_TCINT if tc else 0),
                                  fc32_full_scale=8.0 if tc else 0.0)
        want["stream"] += g0
        sel = got[(got["stream"] >= g0) & (got["stream"] < g0 + 128)]
        assert_recs_equal(sel, want)
    _save("c5_shard_parity_%s_pytest.json" % frontend, {"streams": S, "frontend": frontend, "records": int(len(got)),
                                                        "oracle_seconds": time.time() - t0, "bit_identical_to_oracle": True})
