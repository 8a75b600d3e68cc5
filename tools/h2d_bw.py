#!/usr/bin/env python
"""Pinned H2D / D2H bandwidth with and without binding the process to the GPU's NUMA-local cores.  Diagnostic only."""
import os, sys, glob
import torch

def gpu_local_cpus(index=0):
    p = torch.cuda.get_device_properties(index)
    bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    path = f"/sys/bus/pci/devices/{bdf}/local_cpulist"
    txt = open(path).read().strip()
    cpus = set()
    for part in txt.split(","):
        if "-" in part:
            a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    node = open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip()
    return bdf, cpus, node

def bw(tag):
    for mb in (24, 41):
        h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
        h.zero_()
        d = torch.empty_like(h, device="cuda")
        for _ in range(3): d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): d.copy_(h, non_blocking=True)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print(f"[{tag}] H2D {mb} MiB pinned: {ms:.3f} ms = {(mb << 20) / ms / 1e6:.1f} GB/s")

print("allowed cpus:", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:4], "...", "nodes:", [os.path.basename(p) for p in glob.glob("/sys/devices/system/node/node[0-9]*")])
for n in glob.glob("/sys/devices/system/node/node[0-9]*"):
    print(" ", os.path.basename(n), open(n + "/cpulist").read().strip())
torch.cuda.init()
bdf, cpus, node = gpu_local_cpus(0)
print("gpu0", bdf, "numa node", node, "local cpus", len(cpus), sorted(cpus)[:4], "...")
bw("default affinity")
os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or os.sched_getaffinity(0))
bw("gpu-local cores")
