"""Host-side placement for the end-to-end (host tape) path.

The pipelined host entry point (``IA2CTrainer.train_episodes_host`` -> ``ia2c_train_episodes_host``) runs at the rate of its
one pinned H2D copy per episode.  On a two-socket GPU box that copy runs at ~52 GB/s when the pinned pages live on the
NUMA node the GPU's PCIe root hangs off, and at ~20 GB/s when they live on the other socket (measured, profiles/).  Linux
places pinned pages on the node of the thread that first touches them, so binding the process to the GPU-local CPUs
BEFORE the tapes are allocated is all it takes.  Nothing here touches the device; it is opt-in (a process-wide affinity
change is the caller's decision): bench.py and the launcher call it, a library user calls it once per rank.
"""
from __future__ import annotations

import os


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_cpus(device_index=0):
    """-> (numa_node, set of CPU ids local to the GPU) or (None, None) when the platform does not say."""
    import torch

    p = torch.cuda.get_device_properties(device_index)
    try:
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node >= 0:
            return node, _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
    except Exception:
        pass
    try:   # NVML knows the ideal CPU set even when sysfs reports node -1 (virtualised PCI topology)
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{p.pci_domain_id:08x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0".encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        if cpus and len(cpus) < os.cpu_count():
            return None, cpus
    except Exception:
        pass
    return None, None


def bind_to_gpu_numa_node(device_index=0):
    """Restrict this process to the CPUs local to the GPU (so that pinned host memory allocated afterwards is local too).
    -> dict describing what was done (for logs / the bench record).  Never raises."""
    info = {"bound": False, "numa_node": None, "cpus": None}
    try:
        before = os.sched_getaffinity(0)
        node, cpus = gpu_numa_cpus(device_index)
        info["numa_node"] = node
        if not cpus:
            info["why"] = "GPU-local CPU set unknown (no numa_node in sysfs, no NVML affinity)"
            return info
        usable = cpus & before
        if not usable:
            info["why"] = "GPU-local CPUs are outside this process's allowed set"
            return info
        os.sched_setaffinity(0, usable)
        info.update(bound=True, cpus=len(usable), was=len(before))
    except Exception as exc:   # containers without the syscall, odd topologies
        info["why"] = f"{type(exc).__name__}: {exc}"
    return info


def unbind(cpus=None):
    """Give the process (or a child about to exec) every CPU again."""
    try:
        os.sched_setaffinity(0, cpus if cpus is not None else range(os.cpu_count()))
    except Exception:
        pass
