"""Pin a rank's host threads to the cores next to its GPU (one process per GPU).

The host-pointer batch entry points stream every chunk through pinned staging memory; which NUMA node that memory and
the copying thread live on decides how much of the PCIe link a rank gets when several ranks share one host.  Called
before the first CUDA allocation so that pinned buffers are first-touched on the right node.
"""
from __future__ import annotations

import os


def _gpu_cpus(index: int):
    """Cores the driver reports as local to GPU `index` (NVML), or None."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        return [c for c in cpus if c < n] or None
    except Exception:
        return None


def pin_to_gpu_node(local_rank: int, local_world: int | None = None) -> str | None:
    """Restrict this process to the cores of GPU `local_rank`'s NUMA node; when that is "every core" (one node, or no
    information), to an even share of the allowed cores so that ranks do not migrate over each other.  Returns a short
    description of what was done (None if nothing)."""
    if not hasattr(os, "sched_setaffinity"):
        return None
    if local_world is None:
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
    allowed = sorted(os.sched_getaffinity(0))
    cpus = _gpu_cpus(local_rank)
    how = "gpu-local cores"
    if cpus:
        cpus = [c for c in cpus if c in allowed]
    if not cpus or len(cpus) == len(allowed):
        if local_world <= 1:
            return None
        share = max(1, len(allowed) // local_world)
        cpus = allowed[local_rank * share:(local_rank + 1) * share] or allowed
        how = "even share of the cores"
    try:
        os.sched_setaffinity(0, cpus)
    except OSError:
        return None
    return f"{how}: {cpus[0]}-{cpus[-1]} ({len(cpus)})"
