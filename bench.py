#!/usr/bin/env python
"""bench.py -- PSS+SSS search throughput (BASELINE.json metric) on N B200s of one node.

Workload (config C5's per-GPU shard): 512 concurrent 30.72 Msps fc32 streams per
GPU, decimate-by-16 front end, three-root PSS search + tracking state machine + SSS, fed as
halo-segmented 100 ms segments.  One "step" = one 100 ms segment of all 512 streams
(1.573 G input samples, 12.6 GB fc32 -- larger than L2, so no flush is needed between steps);
the default 10 steps are the config's 1 s per stream.  Streams shard by rank with no
collective on the data path (weak scaling): value = all ranks' samples / max-over-ranks time.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference ...                     # CPU reference arm (oracle, FFT mode)

Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gr-ltetrigger_b200", "python"))

SEARCH_RATE = 1.92e6
F_PSS = 3 * 128 * 8 + 18                     # flop per search-rate sample, direct form


def ntaps(decim):
    """gr-filter rational_resampler_ccc(1, D) default design: tap count."""
    if decim <= 1:
        return 0
    n = int((7.0 / 0.1102 + 8.7) / (22.0 * 0.1 / decim))
    return n + 1 - (n & 1)


FMT_BYTES = {"fc32": 8, "sc16": 4, "sc8": 2}
FP32_PEAK_TFLOPS = 72.4                      # measured: tools/ubench_fp32.cu on this pool's B200 (profiles/ubench_fp32_r01.jsonl)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=512, help="streams per GPU")
    ap.add_argument("--segment-ms", type=int, default=100)
    ap.add_argument("--decim", type=int, default=16)
    ap.add_argument("--format", default="fc32", choices=["fc32", "sc16", "sc8"])
    ap.add_argument("--corr", default="fft", choices=["direct", "fft"],
                    help="matched-filter evaluation: folded direct form or overlap-save FFT blocks")
    ap.add_argument("--snr-db", type=float, default=5.0)
    ap.add_argument("--noise-only", action="store_true",
                    help="no cell in any stream: every chain searches every window (the detector's worst case; "
                         "the default batch carries a cell per stream, so its matching chain skips 8 of 9 searches)")
    ap.add_argument("--cfo-hz", type=float, default=0.0,
                    help="carrier frequency offset of the synthetic streams: stream s gets cfo * (s mod 5 - 2) / 2, i.e. offsets "
                         "between -cfo and +cfo (the tracker's CFO correction then works on non-trivial phase ramps)")
    ap.add_argument("--unique", type=int, default=8, help="distinct synthetic captures tiled over the streams")
    ap.add_argument("--e2e-streams", type=int, default=128)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-streams", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-formats", action="store_true", help="skip the e2e legs on the other wire formats")
    ap.add_argument("--no-e2e-balance", action="store_true",
                    help="N > 1: equal stream shares per rank in the e2e leg instead of shares by measured host-link rate")
    ap.add_argument("--workload", default="c5", choices=["c5", "c1", "c2", "c3"],
                    help="c5 (default): BASELINE's batched shard, the metric's configuration; c1/c2/c3: the single-stream "
                         "configurations (latency to first track, C-ABI and block-path throughput; tools/bench_single.py)")
    ap.add_argument("--frontend", default="auto", choices=["auto", "fp32", "tc"],
                    help="fp32: canonical FFMA2 decimator; tc: exact-integer tensor-core front end (D = 16; fc32 input on a "
                         "23-bit fixed-point grid); auto: tc where it exists (D = 16: it meets the tolerance and wins), else "
                         "fp32.  The other mode is timed on the same input and reported under other_frontend")
    ap.add_argument("--pipeline", default="overlap", choices=["overlap", "serial"],
                    help="overlap: track+SSS of call i under the front end of call i+1 (two streams); serial: one stream")
    ap.add_argument("--no-spot-check", action="store_true", help="skip the per-rank oracle check after timing")
    ap.add_argument("--no-alt", action="store_true", help="skip the run of the other front-end mode on the same input")
    ap.add_argument("--no-alone", action="store_true", help="skip the 3-step serial pass that times every kernel alone")
    ap.add_argument("--sustained-s", type=float, default=2.0, help="length of the extra sustained run (0: skip)")
    a = ap.parse_args()
    if a.frontend == "auto":
        # measured (profiles/tc_rates_r02.log): the tensor-core front end wins wherever it exists except fc32 at D <= 4,
        # where a tile is so few input bytes that its epilogue outweighs the FFMA2 kernel's 129 / 65 taps
        wins = tc_exists(a.format, a.decim) and not (a.format == "fc32" and a.decim <= 4)
        a.frontend = "tc" if (wins and a.workload == "c5" and a.impl == "b200") else "fp32"
    return a


def tc_exists(fmt, decim):
    """(format, rate) pairs the tensor-core front end is built for (include/ltetrigger_b200.h LTB_FRONTEND_TC_INT)."""
    return decim in {"fc32": (2, 4, 8, 12, 16, 24, 32), "sc16": (4, 8, 12, 16, 24, 32), "sc8": (8, 16, 24, 32)}[fmt]


def workload_name(a):
    tag = "C5 shard" if (a.streams, a.decim) == (512, 16) else "C4-like batch" if a.decim == 1 else "batch"
    return "%s: %d streams/GPU x %.2f Msps %s, D=%d, %d ms segments" % (
        tag, a.streams, 1.92 * a.decim, a.format, a.decim, a.segment_ms)


def host_unique(a, n):
    """`unique` seeded synthetic captures of one segment at the input rate (numpy, host)."""
    from ltetrigger_b200 import synth
    perm = np.random.default_rng(2026).permutation(504)
    base = np.empty((a.unique, n), np.complex64)
    for u in range(a.unique):
        # noise-free and frame-periodic (segment = whole frames), so tiling in time stays a valid downlink
        base[u] = synth.capture(int(perm[u]), n, snr_db=None, decim=a.decim, seed=77 + u, offset=0)
    return base, perm[:a.unique]


def host_batch(base, n_streams, snr_db, seed):
    """Tile the unique captures over n_streams with per-stream timing shifts and AWGN (host)."""
    rng = np.random.default_rng(seed)
    u, n = base.shape
    out = np.empty((n_streams, n), np.complex64)
    sigma = np.sqrt(1.0 / (10.0 ** (snr_db / 10.0)) / 2.0)
    for s in range(n_streams):
        shift = int(rng.integers(0, n))
        x = np.roll(base[s % u], shift)
        noise = rng.standard_normal((n, 2), dtype=np.float32)
        out[s] = x + sigma * (noise[:, 0] + 1j * noise[:, 1])
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        self.active, self.ready = False, threading.Event()

    def begin(self):
        """Start recording (call right before the timed region; NVML is already initialised)."""
        self.ready.wait(timeout=10)
        self.samples, self.reasons = [], set()
        self.active = True

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            self.ready.set()
            while not self.stop_flag:
                if self.active:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                time.sleep(0.002)
        except Exception as e:  # NVML missing: report it rather than inventing numbers
            self.reasons.add("nvml_unavailable: %s" % type(e).__name__)
            self.ready.set()

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_reference_run(iq, decim, nthreads=0, conv_fft=True):
    """The reference's algorithm class on the host cores: per window and root one zero-padded
    9728-point FFT convolution (oracle ORC_CONV_FFT), GR-default Kaiser decimator in front."""
    from oracle import oracle as O
    t0 = time.perf_counter()
    recs = O.trigger_run(iq, decim=decim, psr_threshold=4.0, conv_mode=O.CONV_FFT if conv_fft else O.CONV_DIRECT,
                         nthreads=nthreads)
    return time.perf_counter() - t0, recs


def run_reference(a, rank, world):
    if rank != 0:
        return
    n = int(a.segment_ms * 1e-3 * SEARCH_RATE) * a.decim
    base, _ = host_unique(a, n)
    iq = host_batch(base, a.cpu_streams, a.snr_db, seed=5)
    cores = os.cpu_count()
    for _ in range(max(a.warmup, 0) and 1):
        cpu_reference_run(iq[:max(2, cores // 4)], a.decim)
    times = []
    for _ in range(a.steps):
        dt, recs = cpu_reference_run(iq, a.decim)
        times.append(dt)
    total = sum(times)
    value = a.cpu_streams * n * a.steps / total / 1e6
    sample = "%d streams x %d ms per step (same synthetic workload), oracle ORC_CONV_FFT (SIMD decimator, FFT convolution), %d threads" % (
        a.cpu_streams, a.segment_ms, cores)
    line = {
        "impl": "reference", "metric": "PSS+SSS search Msamples/s", "value": value, "unit": "Msamples/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "cpu_sample_streams": a.cpu_streams},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    if a.workload != "c5":
        # single-stream configurations: replicas only beyond one GPU, so rank 0 alone runs them
        if rank == 0:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_single
            r = bench_single.run(a.workload, device=local_rank)
            line = {"metric": "PSS+SSS search Msamples/s", "value": r["c_abi"]["submit_collect_msamples_per_s"], "unit": "Msamples/s",
                    "n_gpus": 1, "steps": 1, "warmup": 3, "ms_per_step": r["c_abi"]["ms_per_100ms_call"], "higher_is_better": True,
                    "scaling": "replicas only", "vs_baseline": None, "dtype": "f32", "data": "reference test frame, tiled",
                    "config": {"workload": "%s: %s @ %.2f Msps, one stream" % (a.workload, r["fixture"], r["sample_rate_msps"])},
                    "e2e": {"value": r["c_abi"]["submit_collect_msamples_per_s"], "unit": "Msamples/s",
                            "api": "ltb_trigger_submit_host + ltb_trigger_collect, host buffers"},
                    "cpu_baseline": {"value": r["cpu_port"]["msamples_per_s"], "unit": "Msamples/s", "cores": r["cpu_port"]["cores"],
                                     "kind": "port", "sample": r["cpu_port"]["what"]},
                    "single_stream": r}
            print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    import ltetrigger_b200 as lt

    if not torch.cuda.is_available() or lt.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # one process per GPU: run on the cores next to that GPU, so the pinned buffers of the
        # end-to-end leg are first touched (and the copy engine reads them) on the GPU's NUMA node
        # (at N = 1 the process keeps every core: the CPU baseline uses them)
        try:
            import pynvml as nv
            nv.nvmlInit()
            nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(local_rank))
            numa = "cores %d" % len(os.sched_getaffinity(0))
        except Exception as e:
            numa = "unpinned (%s)" % type(e).__name__

    fmts = {"fc32": lt.FMT_FC32, "sc16": lt.FMT_SC16, "sc8": lt.FMT_SC8}
    fmt = fmts[a.format]
    bps = FMT_BYTES[a.format]
    n = int(a.segment_ms * 1e-3 * SEARCH_RATE) * a.decim           # input samples per stream per step
    m = n // a.decim

    # ---- synthetic input, resident in HBM before the timed region --------------------------
    base, cell_ids = host_unique(a, n)
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + rank)
    base_d = torch.from_numpy(base).to(dev)
    sigma = float(np.sqrt(1.0 / (10.0 ** (a.snr_db / 10.0)) / 2.0))
    x = torch.empty((a.streams, n), dtype=torch.complex64, device=dev)
    shifts = torch.randint(0, n, (a.streams,), generator=torch.Generator().manual_seed(7 + rank))
    for s in range(a.streams):
        noise = torch.randn((n, 2), generator=g, device=dev, dtype=torch.float32)
        if a.noise_only:
            x[s] = float(np.sqrt(0.5)) * torch.view_as_complex(noise)
        else:
            sig = torch.roll(base_d[s % a.unique], int(shifts[s]))
            if a.cfo_hz:
                f = a.cfo_hz * ((s % 5) - 2) / 2.0 / (SEARCH_RATE * a.decim)
                ph = (2.0 * np.pi * f) * torch.arange(n, device=dev, dtype=torch.float64)
                sig = sig * torch.polar(torch.ones_like(ph), ph).to(torch.complex64)
            x[s] = sig + sigma * torch.view_as_complex(noise)
    del noise, base_d
    def quantise(xc, name):
        """fc32 -> the named wire format (full scale = 8 x the signal's rms)."""
        if name == "fc32":
            return xc
        xr = torch.view_as_real(xc)
        if name == "sc16":
            return torch.clamp(torch.round(xr * (32767.0 / 8.0)), -32768, 32767).to(torch.int16).contiguous()
        return torch.clamp(torch.round(xr * (127.0 / 8.0)), -128, 127).to(torch.int8).contiguous()

    d_in = quantise(x, a.format)
    if a.format != "fc32":
        x_e2e = x[:min(a.e2e_streams, a.streams)].clone()         # fc32 source of the per-format e2e legs
        del x
    else:
        x_e2e = x
    torch.cuda.synchronize()

    stream = torch.cuda.current_stream()
    corr_mode = lt.CORR_FFT if a.corr == "fft" else lt.CORR_DIRECT
    pipeline = lt.PIPE_OVERLAP if a.pipeline == "overlap" else lt.PIPE_SERIAL
    ptr, stride = d_in.data_ptr(), n * bps
    FC32_FULL_SCALE = 8.0      # fc32 through the integer front end: 23-bit fixed point over the range the sc16 / sc8
                               # quantiser above uses (8 x the signal's rms)

    def mode_kw(frontend, format_name):
        """(Trigger keyword arguments, oracle conv_mode flag, oracle full scale) of a front-end mode."""
        if frontend != "tc" or not tc_exists(format_name, a.decim):
            return {}, 0, 0.0
        from oracle import oracle as O_
        kw = {"frontend_mode": lt.FRONTEND_TC_INT}
        if format_name == "fc32":
            kw["fc32_full_scale"] = FC32_FULL_SCALE
        return kw, O_.FRONT_TCINT, kw.get("fc32_full_scale", 0.0)

    names = ["frontend(convert+decimate)", "pss_corr(3 roots)", "pss_track", "sss"]
    names_short = ["frontend", "pss_corr", "pss_track", "sss"]
    hbm_peak = 6535.7
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass

    def timed_run(frontend, pipe, steps, warmup, keep_open=False):
        """W warm-up and K timed steps of one front-end mode: CUDA events on the launch stream around the K steps,
        barrier + synchronize on both sides, max over ranks; per-stage CUDA-event times of the same steps."""
        kw = mode_kw(frontend, a.format)[0]
        trig = lt.Trigger(n_streams=a.streams, decim=a.decim, psr_threshold=4.0, max_chunk=n, input_format=fmt,
                          record_all=False, device=local_rank, cuda_stream=stream.cuda_stream, corr_mode=corr_mode,
                          pipeline=pipe, **kw)
        sampler = ClockSampler(local_rank)
        sampler.start()
        for _ in range(warmup):
            trig.process_device_ptr(ptr, stride, n)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sampler.begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r = {"stage_ms": np.zeros(4), "launches": 0, "n_cells": 0, "last_recs": None}

        def account(recs):
            r["last_recs"] = recs
            r["stage_ms"] += np.array(trig.last_kernel_times())
            r["launches"] += trig.last_timing()[1]
            r["n_cells"] += int(((recs["flags"] & lt.F_CELL) != 0).sum())

        torch.cuda.synchronize()
        e0.record(stream)
        # two calls in flight (ltb_trigger_submit_device / ltb_trigger_collect): call i+1 is enqueued
        # before the records of call i are read, so the stream does not idle on the host round trip
        trig.submit_device_ptr(ptr, stride, n)
        for _ in range(steps - 1):
            trig.submit_device_ptr(ptr, stride, n)
            account(trig.collect())
        account(trig.collect())
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        r["clocks"] = sampler.result()
        if world > 1:
            tt = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.barrier()
            ms = float(tt.item())
        r["elapsed_ms"] = ms
        r["value"] = float(a.streams) * n * steps * world / (ms * 1e-3) / 1e6
        r["stage_ms"] = r["stage_ms"] / steps
        r["last_recs"] = r["last_recs"].copy()
        if keep_open:
            r["trig"] = trig
        else:
            trig.close()
        return r

    def roofline_of(frontend, r, alone_ms=None):
        """Roofline object of the dominant kernel of a mode (live CUDA-event stage times of the timed steps)."""
        stage_ms = r["stage_ms"]
        alg_flop = [4.0 * ntaps(a.decim) * m * a.streams, float(F_PSS) * m * a.streams, 0.0, 0.0]
        dom = int(np.argmax(stage_ms))
        if alg_flop[dom] == 0.0:
            dom = int(np.argmax(stage_ms[:2]))
        achieved = alg_flop[dom] / (stage_ms[dom] * 1e-3) / 1e12
        alg_bytes = [float(bps) * n * a.streams + 8.0 * m * a.streams, 20.0 * m * a.streams, 0.0, 0.0]
        traffic, traffic_src = None, None                           # dram read+write per launch, ncu --set full
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            ent = tr.get(frontend, tr) if frontend in tr else (tr if frontend == "fp32" else {})
            if ent.get("workload") == workload_name(a):
                traffic = ent["dram_bytes_per_launch"].get(names_short[dom])
                traffic_src = ent.get("source")
        except Exception:
            pass
        f_alg = (F_PSS + 4.0 * ntaps(a.decim)) / a.decim            # flop per input sample, SURVEY 8d
        per_gpu_rate = r["value"] * 1e6 / world
        # flop the kernels execute per input sample: the decimator runs the direct form (4 flop per real
        # tap and complex sample); the FFT correlator ~4700 FP32 operations per lane and 896-output block
        # (DESIGN.md K2f), the folded direct form 64 FADD2 + 260 FFMA2 per search-rate sample
        f_corr_exec = (2.0 * 4700 * 32 / 896) if a.corr == "fft" else (64 * 2 + 260 * 4)
        f_exec = (f_corr_exec + 4.0 * ntaps(a.decim)) / a.decim
        hbm_gbs = alg_bytes[dom] / (stage_ms[dom] * 1e-3) / 1e9
        # the integer tensor-core front end does its multiply-adds on the tensor pipe (30 % busy, profiles/): what
        # binds that kernel is memory, so its roofline is HBM; the float32 kernels stay FFMA bound
        tc_dom = frontend == "tc" and dom == 0
        head = ({"bound": "hbm", "kernel": names[dom], "achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak}
                if tc_dom else
                {"bound": "fp32", "kernel": names[dom], "achieved": achieved, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s",
                 "frac": achieved / FP32_PEAK_TFLOPS})
        roofline = dict(head)
        roofline.update({
            "traffic": traffic,
            "traffic_source": traffic_src or "profiles/traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)",
            "algorithmic_bytes_per_launch": alg_bytes[dom],
            "hbm": {"achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)"},
            "fp32": {"achieved": achieved, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP32_PEAK_TFLOPS,
                     "note": "direct-form algorithmic flop of the dominant kernel; with the tc front end they run as int8 tensor-core MACs"},
            "peak_source": ("MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if tc_dom else
                            "measured FFMA peak, tools/ubench_fp32.cu (profiles/ubench_fp32_r01.jsonl); MEASURED_PEAKS.json has no fp32 figure"),
            "stage_ms": {k: float(v) for k, v in zip(names, stage_ms)},
            "stages_overlap": a.pipeline == "overlap",      # track + sss of call i run under the front end of call i+1:
                                                            # the stage times then sum to more than ms_per_step
            "path_frac_of_fp32": f_alg * per_gpu_rate / (FP32_PEAK_TFLOPS * 1e12),
            "path_frac_of_fp32_executed": f_exec * per_gpu_rate / (FP32_PEAK_TFLOPS * 1e12),
            "flop_per_input_sample": {"algorithmic_direct_form": f_alg, "executed": f_exec},
            "path_frac_of_hbm": bps * per_gpu_rate / (hbm_peak * 1e9),
            "note": "achieved = SURVEY 8d algorithmic bytes (hbm) or direct-form flop (fp32) of the dominant kernel per launch / its CUDA-event duration inside the timed (pipelined) steps; kernel_alone repeats it with the stages run back to back on one stream; path_frac_of_fp32 uses the same direct-form count and exceeds 1 because the correlator runs as FFT blocks (or folds the taps); path_frac_of_fp32_executed counts the flop the kernels execute",
        })
        if alone_ms is not None:
            ka = alone_ms[dom]
            roofline["kernel_alone"] = {
                "ms": float(ka), "stage_ms": {k: float(v) for k, v in zip(names, alone_ms)},
                "frac": (alg_bytes[dom] / (ka * 1e-3) / 1e9 / hbm_peak) if tc_dom else (alg_flop[dom] / (ka * 1e-3) / 1e12 / FP32_PEAK_TFLOPS),
                "what": "the same kernels with the stages of a call back to back on one stream (pipeline serial, 3 steps): no other kernel shares the SMs"}
        return roofline

    main = timed_run(a.frontend, pipeline, a.steps, a.warmup, keep_open=True)
    trig = main["trig"]
    elapsed_ms, value, launches, n_cells, clocks = main["elapsed_ms"], main["value"], main["launches"], main["n_cells"], main["clocks"]
    last_recs = main["last_recs"]
    frontend_kw, oracle_front_flag, oracle_fs = mode_kw(a.frontend, a.format)
    alone = None
    if a.pipeline == "overlap" and not a.no_alone:
        alone = timed_run(a.frontend, lt.PIPE_SERIAL, 3, 2)["stage_ms"]
    roofline = roofline_of(a.frontend, main, alone)

    frontend_text = {"fp32": "fp32 (canonical float32 expression trees, FFMA2)",
                     "tc": "tc (exact integer arithmetic on the tensor cores: tcgen05.mma kind::i8, TMA, TMEM%s)" % (
                         "; fc32 samples taken as 23-bit fixed point over +-%.1f" % FC32_FULL_SCALE if a.format == "fc32" else "")}
    out = {
        "metric": "PSS+SSS search Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": elapsed_ms / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if a.frontend == "fp32" else "s8 x s8 -> s32 (front end, exact), f32 (correlator, tracker, SSS)",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "streams_per_gpu": a.streams, "decim": a.decim,
                   "format": a.format, "frontend": frontend_text[a.frontend], "correlator": a.corr, "pipeline": a.pipeline, "segment_ms": a.segment_ms, "snr_db": a.snr_db,
                   "l2": "inputs larger than L2 (%.1f GB per step)" % (a.streams * n * bps / 1e9),
                   "input": ("noise only (no cell in any stream)" if a.noise_only else
                             "%d seeded synthetic LTE captures tiled over the streams with per-stream timing shift + AWGN%s" % (
                                 a.unique, (", carrier offsets up to +-%.0f Hz" % a.cfo_hz) if a.cfo_hz else "")),
                   "cells_tagged_per_step": n_cells / a.steps},
        "clocks": clocks, "gpu_launches": launches, "roofline": roofline,
    }

    # ---- the other front-end mode on the same input, same steps (D = 16 only: the tensor-core kernel's rate) ----
    if tc_exists(a.format, a.decim) and not a.no_alt:
        other = "fp32" if a.frontend == "tc" else "tc"
        trig.close()
        alt = timed_run(other, pipeline, a.steps, a.warmup)
        alt_alone = timed_run(other, lt.PIPE_SERIAL, 3, 2)["stage_ms"] if (a.pipeline == "overlap" and not a.no_alone) else None
        out["other_frontend"] = {"frontend": frontend_text[other], "value": alt["value"], "unit": "Msamples/s",
                                 "ms_per_step": alt["elapsed_ms"] / a.steps, "steps": a.steps, "warmup": a.warmup,
                                 "clocks": alt["clocks"], "gpu_launches": alt["launches"],
                                 "cells_tagged_per_step": alt["n_cells"] / a.steps,
                                 "roofline": roofline_of(other, alt, alt_alone)}
        trig = lt.Trigger(n_streams=a.streams, decim=a.decim, psr_threshold=4.0, max_chunk=n, input_format=fmt,
                          record_all=False, device=local_rank, cuda_stream=stream.cuda_stream, corr_mode=corr_mode,
                          pipeline=pipeline, **frontend_kw)
        for _ in range(2):
            trig.process_device_ptr(ptr, stride, n)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # ---- the same step sustained for >= 2 s: the clock the board settles at under its power cap ----
    if a.sustained_s > 0:
        k_sus = max(a.steps, int(np.ceil(a.sustained_s * 1e3 / (elapsed_ms / a.steps))))
        s2 = ClockSampler(local_rank)
        s2.start()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        s2.begin()
        e0.record(stream)
        trig.submit_device_ptr(ptr, stride, n)
        for _ in range(k_sus - 1):
            trig.submit_device_ptr(ptr, stride, n)
            trig.collect()
        trig.collect()
        e1.record(stream)
        torch.cuda.synchronize()
        sus_ms = e0.elapsed_time(e1)
        c2 = s2.result()
        if world > 1:
            tt = torch.tensor([sus_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            sus_ms = float(tt.item())
        out["sustained"] = {"value": float(a.streams) * n * k_sus * world / (sus_ms * 1e-3) / 1e6, "unit": "Msamples/s",
                            "steps": k_sus, "seconds": sus_ms * 1e-3, "ms_per_step": sus_ms / k_sus, "clocks": c2}

    # ---- every rank checks records of its own shard against the oracle (the checker, not the product) ----
    # the benchmarked configuration (rate, format, correlator, front-end mode) on eight of this rank's own
    # streams, every record compared bit for bit; the verdicts are AND-reduced over the ranks
    if not a.no_spot_check:
        from oracle import oracle as O
        nchk = min(8, a.streams)
        pick = torch.linspace(0, a.streams - 1, nchk).round().long().to(dev)
        iq = d_in.index_select(0, pick).cpu().numpy()
        chk = lt.Trigger(n_streams=nchk, decim=a.decim, psr_threshold=4.0, max_chunk=n, input_format=fmt,
                         device=local_rank, corr_mode=corr_mode, **frontend_kw)
        got = chk.run(iq)
        chk.close()
        conv = (O.CONV_OS if a.corr == "fft" else O.CONV_DIRECT) | oracle_front_flag
        want = O.trigger_run(iq, decim=a.decim, fmt=fmt, psr_threshold=4.0, conv_mode=conv, fc32_full_scale=oracle_fs)
        same = len(got) == len(want)
        for f in (want.dtype.names if same else ()):
            g_, w_ = got[f], want[f]
            if g_.dtype.kind == "f":
                same = same and bool(((g_.view(np.uint32) == w_.view(np.uint32)) | ((g_ == 0) & (w_ == 0))).all())
            else:
                same = same and bool((g_ == w_).all())
        # the two front ends against each other on the same streams (north_star: decisions bit-exact, correlation
        # magnitudes and PSR within 1e-4 relative): the integer front end's records next to the canonical float32 ones
        agree, worst = 1, 0.0
        if tc_exists(a.format, a.decim):
            okw = mode_kw("fp32" if a.frontend == "tc" else "tc", a.format)[0]
            chk = lt.Trigger(n_streams=nchk, decim=a.decim, psr_threshold=4.0, max_chunk=n, input_format=fmt,
                             device=local_rank, corr_mode=corr_mode, **okw)
            oth = chk.run(iq)
            chk.close()
            keys = ("stream", "n_id_2", "win_start", "emit_start", "flags", "peak_pos", "m0", "m1", "n_id_1", "cell_id")
            agree = int(len(oth) == len(got) and all(bool((oth[f] == got[f]).all()) for f in keys))
            if not agree and len(oth) == len(got) and all(bool((oth[f] == got[f]).all()) for f in keys if f != "peak_pos"):
                # a window under the threshold is noise: its argmax may sit on another of two nearly equal maxima
                over = ((got["flags"] | oth["flags"]) & lt.F_OVER) != 0
                agree = int(bool((oth["peak_pos"][over] == got["peak_pos"][over]).all()))
            if agree:
                for f in ("psr", "peak_value"):
                    den = np.maximum(np.abs(got[f]), 1e-30)
                    fin = np.isfinite(got[f]) & np.isfinite(oth[f])
                    if fin.any():
                        worst = max(worst, float((np.abs(got[f] - oth[f])[fin] / den[fin]).max()))
        verdict = torch.tensor([1 if same else 0, len(want), agree], device=dev, dtype=torch.int64)
        wt = torch.tensor([worst], device=dev, dtype=torch.float64)
        if world > 1:
            v0 = verdict.clone()
            dist.all_reduce(verdict[0:1], op=dist.ReduceOp.MIN)
            dist.all_reduce(verdict[2:3], op=dist.ReduceOp.MIN)
            dist.all_reduce(v0[1:2], op=dist.ReduceOp.SUM)
            dist.all_reduce(wt, op=dist.ReduceOp.MAX)
            verdict[1] = v0[1]
        out["parity_spot_check"] = {"ranks": world, "streams_per_rank": nchk, "records": int(verdict[1].item()),
                                    "bit_identical_to_oracle": bool(verdict[0].item() == 1),
                                    "checker": "oracle/ (CPU restatement), same rate/format/correlator/front end as the timed run"}
        if tc_exists(a.format, a.decim):
            out["tc_vs_fp32"] = {"decisions_identical": bool(verdict[2].item() == 1), "max_rel_diff_psr_peak": float(wt.item()),
                                 "tolerance": 1e-4, "records": int(verdict[1].item()),
                                 "what": "window records of the integer tensor-core front end against the canonical float32 front end on the "
                                         "spot-check streams: window starts, flags, peak positions, m0/m1, N_id_1, cell ids equal; PSR and "
                                         "peak value compared relatively"}

    # ---- host merge of the (tiny) record lists of the last step: the only exchange between ranks ----
    from ltetrigger_b200 import shard
    owned = np.arange(rank * a.streams, (rank + 1) * a.streams, dtype=np.int64)     # weak scaling: rank-major ids
    t0 = time.perf_counter()
    merged = shard.merge_records(shard.to_global(last_recs, owned), dst=0)
    merge_ms = 1e3 * (time.perf_counter() - t0)
    if rank == 0:
        out["merge"] = {"records": int(len(merged)), "ms": merge_ms, "ranks": world,
                        "what": "shard.merge_records of the last step's window records over %s (all-gather, sorted by stream/root/window)"
                                % ("NCCL" if world > 1 else "no process group: local sort")}

    # ---- end to end through the C ABI with host buffers (H2D + D2H inside the timed region) --
    if not a.no_e2e:
        se = min(a.e2e_streams, a.streams)
        trig.close()
        # Several GPUs of one box do not get equal shares of the host side (profiles/e2e_probe_r02_n8.json: with eight
        # copies at once four links run at 23.8 and four at 35.5 GB/s, each 55.6 GB/s alone), and the end-to-end leg is as
        # slow as its slowest rank.  Streams are independent, so the host deals them in proportion to what each rank's link
        # delivers while all ranks copy: measured here, barrier-aligned, before the timed region.
        balance = None
        if world > 1:
            probe = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
            probe.fill_(1)
            pd = torch.empty_like(probe, device=dev)
            pd.copy_(probe, non_blocking=True)
            torch.cuda.synchronize()
            dist.barrier()
            e0.record(stream)
            for _ in range(6):
                pd.copy_(probe, non_blocking=True)
            e1.record(stream)
            torch.cuda.synchronize()
            bw = torch.zeros(world, device=dev, dtype=torch.float64)
            bw[rank] = 6 * probe.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9
            dist.all_reduce(bw, op=dist.ReduceOp.SUM)
            del probe, pd
            bw_l = [float(v) for v in bw.tolist()]
            if not a.no_e2e_balance:
                share = [max(8, min(a.streams, int(round(se * world * v / sum(bw_l))))) for v in bw_l]
                balance = {"concurrent_h2d_gbs": bw_l, "streams_per_rank": share,
                           "what": "streams dealt in proportion to each rank's host-link rate with all ranks copying at once"}
                se = share[rank]
            else:
                balance = {"concurrent_h2d_gbs": bw_l, "streams_per_rank": [se] * world, "what": "equal shares (--no-e2e-balance)"}

        def run_e2e(name):
            src = d_in[:se] if name == a.format else quantise(x_e2e[:se], name)
            host = torch.empty(tuple(src.shape), dtype=src.dtype, pin_memory=True)
            host.copy_(src)
            del src
            torch.cuda.synchronize()
            b = FMT_BYTES[name]
            trig2 = lt.Trigger(n_streams=se, decim=a.decim, psr_threshold=4.0, max_chunk=n, input_format=fmts[name],
                               record_all=False, device=local_rank, cuda_stream=stream.cuda_stream, corr_mode=corr_mode,
                               pipeline=pipeline, **mode_kw(a.frontend, name)[0])
            hptr, hstride = host.data_ptr(), n * b
            for _ in range(3):                                      # warm-up: allocates both staging buffers,
                trig2.submit_host_ptr(hptr, hstride, n)             # touches every pinned page
                trig2.submit_host_ptr(hptr, hstride, n)
                trig2.collect(); trig2.collect()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0.record(stream)
            d2h = 0
            trig2.submit_host_ptr(hptr, hstride, n)             # two calls in flight: the H2D of call i+1
            for _ in range(a.e2e_steps - 1):                     # overlaps the kernels of call i
                trig2.submit_host_ptr(hptr, hstride, n)
                d2h += trig2.collect().nbytes
            d2h += trig2.collect().nbytes
            e1.record(stream)
            torch.cuda.synchronize()
            e2e_ms = e0.elapsed_time(e1)
            e2e_ms_local = e2e_ms
            if world > 1:
                tt = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                e2e_ms = float(tt.item())
            trig2.close()
            # the host link alone: the same pinned buffer copied to the device, nothing else running
            dst = torch.empty_like(host, device=dev)
            dst.copy_(host, non_blocking=True)
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(4):
                dst.copy_(host, non_blocking=True)
            e1.record(stream)
            torch.cuda.synchronize()
            h2d_gbs = 4 * host.numel() * host.element_size() / (e0.elapsed_time(e1) * 1e-3) / 1e9
            del dst
            link = {("h2d_copy_alone_gbs" if world == 1 else "h2d_copy_all_ranks_at_once_gbs"): h2d_gbs,
                    "frac_of_h2d_copy": (se * n * b * a.e2e_steps / (e2e_ms_local * 1e-3) / 1e9) / h2d_gbs}
            tot = torch.tensor([float(se), float(d2h // a.e2e_steps)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            se_all = float(tot[0].item())
            res = {"value": se_all * n * a.e2e_steps / (e2e_ms * 1e-3) / 1e6, "unit": "Msamples/s",
                   "h2d_bytes_per_step": int(se_all * n * b), "d2h_bytes_per_step": int(tot[1].item()),
                   "streams": int(se_all), "steps": a.e2e_steps, "host_memory": "pinned", "format": name, "cpu_affinity": numa, "host_link": link,
                   "api": "ltb_trigger_submit_host + ltb_trigger_collect (two calls in flight)"}
            if balance is not None:
                res["balance"] = balance
            return host, res

        host, out["e2e"] = run_e2e(a.format)
        # the same streams on the narrower wire formats an SDR delivers: the host link is the e2e bound
        out["e2e_by_format"] = {a.format: out["e2e"]["value"]}
        for name in ("fc32", "sc16", "sc8"):
            if name != a.format and not a.no_e2e_formats:
                out["e2e_by_format"][name] = run_e2e(name)[1]["value"]
        # ---- CPU baseline on the same host sample (rank 0, N=1 only) ---------------------------
        if rank == 0 and world == 1 and not a.no_cpu_baseline and fmt == lt.FMT_FC32:
            sc = min(a.cpu_streams, se)
            iq = host[:sc].numpy()
            cpu_reference_run(iq[:2], a.decim)                     # page in the oracle, build its tables
            dt, reps = 0.0, 0
            while dt < 12.0 and reps < 200:                        # about 12 s of CPU work on all cores
                dt += cpu_reference_run(iq, a.decim)[0]
                reps += 1
            out["cpu_baseline"] = {"value": sc * n * reps / dt / 1e6, "unit": "Msamples/s", "cores": os.cpu_count(),
                                   "kind": "port",
                                   "sample": "%d passes over %d streams x %d ms of the same workload (%.1f s wall); oracle in "
                                             "reference-class mode (SIMD dot-product decimator, 9728-point FFT convolution per window "
                                             "and root), one job per stream / (stream, root) on all cores" % (reps, sc, a.segment_ms, dt)}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
