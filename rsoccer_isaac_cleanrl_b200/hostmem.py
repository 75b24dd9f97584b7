"""Pinned host buffers for the host-facing step (`_FusedView.step_host`).

`pinned_empty` allocates page-locked memory while the calling thread is restricted to the CPUs NVML
reports as local to the GPU: cudaHostAlloc places (first-touches) the pages from the calling thread, so
under the kernel's default local policy they land on the GPU's own NUMA node and the PCIe traffic of
one rank does not cross the socket interconnect. On a single-node host (or without NVML) it is a plain
pinned allocation. Under unified addressing pinned memory is also device-mapped: its `data_ptr()` can be
handed to a kernel as an output pointer (`vss_set_step_packed`).
"""
import os

import torch


def _local_cpus(device):
    """CPUs local to `device` according to NVML, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            bus = torch.cuda.get_device_properties(device).pci_bus_id
            dom = torch.cuda.get_device_properties(device).pci_domain_id
            dev = torch.cuda.get_device_properties(device).pci_device_id
            h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{dom:08x}:{bus:02x}:{dev:02x}.0".encode())
            words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
            cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
            return cpus or None
        finally:
            pynvml.nvmlShutdown()
    except Exception:
        return None


def pinned_empty(shape, dtype, device):
    """Page-locked (and device-mapped) host tensor, placed next to `device` when the host has more
    than one NUMA node."""
    old = None
    try:
        cpus = _local_cpus(torch.device(device))
        allowed = os.sched_getaffinity(0)
        if cpus and (cpus & allowed) and (cpus & allowed) != allowed:
            old = allowed
            os.sched_setaffinity(0, cpus & allowed)
    except Exception:
        old = None
    try:
        t = torch.empty(shape, dtype=dtype).pin_memory()
        t.zero_()  # touch every page from this thread
    finally:
        if old is not None:
            os.sched_setaffinity(0, old)
    return t
