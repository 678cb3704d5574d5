#!/usr/bin/env python
"""Turn ncu outputs into the small text summaries committed under profiles/.

  ncu_summary.py launches <launches.csv> [--last-steps N]     per-launch list of this repo's kernels + shares
  ncu_summary.py raw <report.ncu-rep>                         key metrics per profiled launch (needs ncu on PATH)
  ncu_summary.py hot <report.ncu-rep> [N]                     the N most sampled SASS instructions + code regions by
                                                              executed-instruction count (needs -lineinfo + --import-source)
"""
import collections
import csv
import subprocess
import sys

KEY_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_bytes.sum", "launch__grid_size", "launch__block_size",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg", "sm__inst_executed.sum",
    "smsp__inst_executed.avg.per_cycle_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    ours = ("ltb::", "decimate_", "pss_corr", "pss_track", "sss_kernel", "sss_block", "tail_kernel", "ingest_kernel", "chain_order")
    mine = [r for r in rows[1:] if any(k in r[ci["Kernel Name"]] for k in ours)]
    print("# our kernels in %s: %d launches (torch data-generation kernels of bench.py omitted)" % (path, len(mine)))
    print("id,kernel,grid,block,ns")
    tot = collections.OrderedDict()
    for r in mine:
        name = r[ci["Kernel Name"]].split("(")[0].replace("void ", "")
        ns = float(r[ci["Metric Value"]].replace(",", ""))
        print("%s,%s,%s,%s,%.0f" % (r[ci["ID"]], name, r[ci["Grid Size"]].replace(",", " "), r[ci["Block Size"]].replace(",", " "), ns))
        a = tot.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    s = sum(v[1] for v in tot.values())
    print("# kernel,launches,avg_us,share_of_our_gpu_time")
    for k, (n, t) in tot.items():
        print("# %s,%d,%.1f,%.3f" % (k, n, t / n / 1e3, t / s))


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("kernel: %s" % r[ci["Kernel Name"]].split("(")[0])
        for m in KEY_METRICS:
            if m in ci:
                print("  %-62s %s %s" % (m, r[ci[m]], units[ci[m]]))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        tot = sum(float(r[ci[m]]) * scale[units[ci[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        print("  %-62s %.4f Gbyte" % ("traffic = dram read + write", tot / 1e9))


def hot(path, n=24):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    print("kernel: %s" % rows[0][1])
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    data = []
    for idx, r in enumerate(rows[2:]):
        try:
            data.append((idx, float(r[ci["# Samples"]]), float(r[ci["Instructions Executed"]]), r[ci["Source"]].strip()))
        except (ValueError, IndexError):
            continue
    tot, tote = sum(d[1] for d in data), sum(d[2] for d in data)
    print("stall samples %d, warp instructions executed %d" % (tot, tote))
    print("# code regions (runs of instructions with the same execution count), share of executed instructions / of samples")
    seg, cur = [], None
    for idx, v, e, src in data:
        if cur and abs(cur["e"] - e) <= 0.15 * max(cur["e"], 1):
            cur["n"] += 1; cur["s"] += v; cur["etot"] += e; cur["last"] = idx
        else:
            cur = {"first": idx, "last": idx, "n": 1, "s": v, "e": e, "etot": e, "src": src}
            seg.append(cur)
    for c in sorted(sorted(seg, key=lambda c: -c["etot"])[:12], key=lambda c: c["first"]):
        print("  instr %4d-%4d (%3d)  exec each %10.0f  executed %5.1f%%  samples %5.1f%%  starts: %s" % (
            c["first"], c["last"], c["n"], c["e"], 100 * c["etot"] / tote, 100 * c["s"] / tot, c["src"][:70]))
    print("# most sampled instructions")
    for idx, v, e, src in sorted(sorted(data, key=lambda x: -x[1])[:n]):
        print("  #%4d %5.1f%%  exec %10.0f  %s" % (idx, 100 * v / tot, e, src[:100]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "hot":
        hot(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 24)
    else:
        raw(sys.argv[2])
