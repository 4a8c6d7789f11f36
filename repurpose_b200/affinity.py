"""Host placement for multi-GPU runs: bind a rank's threads (and therefore its first-touch pinned
allocations) to the NUMA node its GPU hangs off.  With eight ranks streaming fp32 features at tens of
GB/s each, pinned buffers on the wrong socket turn the inter-socket link into the bottleneck of the
host->device path.  Best effort: any failure (no sysfs, no NVML, cpuset restrictions) leaves the
process untouched."""
from __future__ import annotations

import os


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device_index: int) -> int | None:
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(visible.split(",")[device_index]) if visible and visible.split(",")[device_index].isdigit() else device_index
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:      # NVML prints an 8-digit domain, sysfs uses 4
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa(device_index: int) -> dict:
    """Restrict this process to the CPUs of the GPU's NUMA node.  Returns what was done (for logs)."""
    info = {"numa_node": None, "cpus": None}
    node = gpu_numa_node(device_index)
    if node is None:
        return info
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(numa_node=node, cpus=len(allowed))
    except Exception:
        pass
    return info
