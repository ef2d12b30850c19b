"""Pinned host buffers on the NUMA node of the GPU that will read them.

With one process per GPU and eight GPUs behind two CPU sockets, pinned staging buffers that all sit on one socket make
every host<->device copy of the far GPUs cross the socket interconnect (measured in round 1: the end-to-end rate per GPU
halves at N >= 4).  Linux places the pages of a fresh allocation on the node of the CPU that touches them first, so the
buffers are allocated (and pinned) while the calling thread is confined to the CPUs of the GPU's node.
"""
from __future__ import annotations

import contextlib
import os
from typing import Optional, Set

import torch


def _parse_cpulist(text: str) -> Set[int]:
    cpus: Set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device_index: int) -> Optional[int]:
    """NUMA node of a CUDA device from sysfs, or None when the platform does not say."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[device_index]) if vis else device_index
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}/numa_node"
        node = int(open(path).read().strip())
        return node if node >= 0 else None
    except Exception:  # noqa: BLE001
        return None


def gpu_cpus(device_index: int) -> Set[int]:
    """CPUs NVML calls local to the GPU (nvmlDeviceGetCpuAffinity): works where sysfs reports no NUMA node for the PCI
    device (virtualised hosts) but the driver still knows the topology.  Empty when NVML does not say either."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[device_index]) if vis else device_index
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys), (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        return cpus if 0 < len(cpus) < n_cpu else set()        # "every CPU" carries no information
    except Exception:  # noqa: BLE001
        return set()


def node_cpus(node: int) -> Set[int]:
    try:
        return _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
    except Exception:  # noqa: BLE001
        return set()


@contextlib.contextmanager
def on_gpu_node(device_index: int):
    """Confine the calling thread to the CPUs of the GPU's NUMA node for the duration of the block (no-op when unknown)."""
    node = gpu_numa_node(device_index)
    cpus = node_cpus(node) if node is not None else gpu_cpus(device_index)
    old = None
    try:
        if cpus and hasattr(os, "sched_getaffinity"):
            old = os.sched_getaffinity(0)
            allowed = cpus & old
            if allowed:
                os.sched_setaffinity(0, allowed)
            else:
                old = None
        yield node
    finally:
        if old is not None:
            os.sched_setaffinity(0, old)


def pinned_like(t: torch.Tensor, device_index: int) -> torch.Tensor:
    """A pinned host copy of ``t`` whose pages live on the NUMA node of GPU ``device_index``."""
    with on_gpu_node(device_index):
        host = torch.empty(t.shape, dtype=t.dtype).pin_memory()    # pages are faulted in (and pinned) by this thread
        host.copy_(t)
    return host


def pinned_empty(shape, dtype, device_index: int) -> torch.Tensor:
    with on_gpu_node(device_index):
        host = torch.empty(shape, dtype=dtype).pin_memory()
        host.zero_()
    return host
