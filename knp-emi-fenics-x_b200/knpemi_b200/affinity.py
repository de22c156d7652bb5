"""Bind the calling process to the CPUs next to a GPU (NUMA locality of pinned buffers).

With several GPUs on a two-socket host, page-locked staging memory that lives on the far
socket makes every host<->device copy cross the inter-socket link.  Pinning the process to
the GPU's local CPUs *before* the buffers are allocated lets first-touch placement put them
on the right node.  Uses NVML (nvidia-ml-py) when present; silently does nothing otherwise.
"""
from __future__ import annotations

import os


def bind_to_device(dev: int) -> list[int]:
    """Restrict this process to the CPU set NVML reports as local to GPU `dev`.
    Returns the CPU list (empty if nothing was changed)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(dev))
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, n_words)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return allowed
    except Exception:
        pass
    return []
