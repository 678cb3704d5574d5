#!/usr/bin/env python
"""What limits the end-to-end (host-buffer) path when N GPUs of one box are fed at once: measured, not assumed.

Run under torchrun with N ranks (one per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/e2e_scaling_probe.py --gib 1 --reps 6

For every rank it measures pinned-host -> device copies of the same size
  alone       one rank copies while the others wait at a barrier (the link of that GPU by itself)
  concurrent  all ranks copy at once, barrier-aligned (what bench.py's end-to-end leg does N times in parallel)
in four variants of how the pinned buffer was made: {default, write-combined} x {process pinned to the cores
next to its GPU (first touch on that NUMA node), unpinned}.  It also records where things are: the NUMA node
of every GPU (sysfs), the node(s) the pinned pages ended up on (/proc/self/numa_maps), the cores each rank
may run on.  Rank 0 prints one JSON line; `sum_concurrent_gbs` against `sum_alone_gbs` says whether the host
side (memory / root complex / hypervisor) rather than the per-GPU link is the limit.
"""
import argparse
import ctypes
import json
import os
import re

import numpy as np
import torch
import torch.distributed as dist


def gpu_numa_node(index):
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        bus = nv.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            return int(f.read().strip()), bus
    except Exception as e:
        return None, "unknown (%s)" % type(e).__name__


def numa_nodes_of(ptr, nbytes):
    """Which NUMA nodes hold the pages of [ptr, ptr+nbytes): parsed from /proc/self/numa_maps."""
    nodes = {}
    try:
        with open("/proc/self/numa_maps") as f:
            for line in f:
                parts = line.split()
                start = int(parts[0], 16)
                if start <= ptr < start + (1 << 40):
                    found = {int(m.group(1)): int(m.group(2)) for m in re.finditer(r"N(\d+)=(\d+)", line)}
                    if start == ptr or (found and start <= ptr):
                        if start == ptr:
                            return found
                        nodes = found
    except Exception:
        return None
    return nodes or None


class HostBuf:
    """nbytes of page-locked host memory, default or write-combined (cudaHostAlloc through libcudart)."""

    def __init__(self, nbytes, write_combined):
        self.nbytes, self.wc = nbytes, write_combined
        if not write_combined:
            self.t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            self.t.fill_(1)                                   # first touch here, on this process's cores
            self.ptr = self.t.data_ptr()
            self._rt = None
        else:
            self._rt = ctypes.CDLL("libcudart.so.12")
            p = ctypes.c_void_p()
            rc = self._rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(4))   # cudaHostAllocWriteCombined
            if rc != 0:
                raise RuntimeError("cudaHostAlloc(write-combined) -> %d" % rc)
            self.ptr = p.value
            ctypes.memset(self.ptr, 1, nbytes)

    def free(self):
        if self._rt is not None:
            self._rt.cudaFreeHost(ctypes.c_void_p(self.ptr))
        else:
            del self.t


def copy_gbs(dst, buf, reps, stream):
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rt.cudaMemcpyAsync(dst.data_ptr(), buf.ptr, buf.nbytes, 1, ctypes.c_void_p(stream.cuda_stream))
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        rt.cudaMemcpyAsync(dst.data_ptr(), buf.ptr, buf.nbytes, 1, ctypes.c_void_p(stream.cuda_stream))
    e1.record(stream)
    torch.cuda.synchronize()
    return reps * buf.nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=6)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = int(a.gib * (1 << 30))
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    all_cores = sorted(os.sched_getaffinity(0))
    node, bus = gpu_numa_node(local)
    info = {"rank": rank, "gpu_pci": bus, "gpu_numa_node": node, "cores_available": len(all_cores)}
    results = {}
    for pin in (False, True):
        if pin:
            try:
                import pynvml as nv
                nv.nvmlInit()
                nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(local))
                info["cores_next_to_gpu"] = len(os.sched_getaffinity(0))
            except Exception as e:
                info["cores_next_to_gpu"] = "unavailable (%s)" % type(e).__name__
        else:
            os.sched_setaffinity(0, all_cores)
        for wc in (False, True):
            key = "%s_%s" % ("wc" if wc else "default", "affine" if pin else "unpinned")
            try:
                buf = HostBuf(nbytes, wc)
            except Exception as e:
                results[key] = {"error": str(e)}
                continue
            placed = numa_nodes_of(buf.ptr, nbytes)
            alone = 0.0
            for r in range(world):                           # one rank at a time
                if world > 1:
                    dist.barrier()
                if r == rank:
                    alone = copy_gbs(dst, buf, a.reps, stream)
            if world > 1:
                dist.barrier()
            conc = copy_gbs(dst, buf, a.reps, stream)        # everybody at once
            if world > 1:
                dist.barrier()
            results[key] = {"alone_gbs": alone, "concurrent_gbs": conc, "pages_on_nodes": placed}
            buf.free()
    mine = {"info": info, "results": results}
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        dist.barrier()
        dist.destroy_process_group()
    else:
        gathered = [mine]
    if rank == 0:
        summary = {}
        for key in gathered[0]["results"]:
            rows = [g["results"][key] for g in gathered]
            if any("error" in r for r in rows):
                summary[key] = {"error": [r.get("error") for r in rows]}
                continue
            summary[key] = {"alone_gbs": [round(r["alone_gbs"], 1) for r in rows],
                            "concurrent_gbs": [round(r["concurrent_gbs"], 1) for r in rows],
                            "sum_alone_gbs": round(sum(r["alone_gbs"] for r in rows), 1),
                            "sum_concurrent_gbs": round(sum(r["concurrent_gbs"] for r in rows), 1),
                            "pages_on_nodes": [r["pages_on_nodes"] for r in rows]}
        host = {"cpu_count": os.cpu_count()}
        try:
            host["numa_nodes"] = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
            with open("/proc/meminfo") as f:
                host["mem_total_gb"] = round(int(f.readline().split()[1]) / 1e6, 1)
            with open("/proc/cpuinfo") as f:
                for line in f:
                    if line.startswith("model name"):
                        host["cpu_model"] = line.split(":", 1)[1].strip()
                        break
            try:
                with open("/sys/hypervisor/type") as f:
                    host["hypervisor"] = f.read().strip()
            except Exception:
                with open("/proc/cpuinfo") as f:
                    host["hypervisor_flag"] = any(" hypervisor" in line for line in f if line.startswith("flags"))
        except Exception as e:
            host["error"] = str(e)
        print(json.dumps({"probe": "h2d_scaling", "n_gpus": world, "gib_per_copy": a.gib, "reps": a.reps, "host": host,
                          "ranks": [g["info"] for g in gathered], "summary": summary}), flush=True)


if __name__ == "__main__":
    main()
