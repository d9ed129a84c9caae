"""Host-side placement for multi-GPU runs: bind the calling thread to the CPUs that are local to a GPU (same NUMA node / PCIe root),
so that its page-locked result buffers are allocated in that node's memory and the device->host copy of the waveforms (the bulk of the
traffic: 4 bytes x 44 100 per audio-second) does not cross the socket interconnect. Eight ranks copying 200 MB each per pass is where
the end-to-end leg lost against the device-resident one at N = 8. Best effort: silently does nothing when NVML or the
affinity information is not available."""
from __future__ import annotations

import os
from typing import List, Optional


def cpus_of_gpu(index: int) -> Optional[List[int]]:
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(index))
        words = (max(os.cpu_count() or 1, 1) + 63) // 64
        masks = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * i + b for i, m in enumerate(masks) for b in range(64) if (int(m) >> b) & 1]
        return cpus or None
    except Exception:           # noqa: BLE001 (no NVML, no permission, index out of range ...)
        return None


def bind_to_gpu(index: int) -> Optional[List[int]]:
    """Restrict the CALLING THREAD to the CPUs local to GPU `index` (intersected with what the process may use).
    Returns the CPU list it bound to, or None when nothing was changed."""
    cpus = cpus_of_gpu(index)
    if not cpus or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        allowed = os.sched_getaffinity(0)
        use = sorted(set(cpus) & set(allowed))
        if not use or len(use) == len(allowed):
            return None
        os.sched_setaffinity(0, use)
        return use
    except OSError:
        return None
