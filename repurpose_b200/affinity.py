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


def _pci_bus_numbers() -> list[int] | None:
    """PCI bus number of every visible GPU in CUDA order, or None when NVML is not usable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        n = pynvml.nvmlDeviceGetCount()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        order = [int(x) for x in visible.split(",")] if visible and all(x.strip().isdigit() for x in visible.split(",")) \
            else list(range(n))
        return [int(pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(i)).bus) for i in order]
    except Exception:
        return None


def pick_device(local_rank: int, world_size: int, buses: list[int] | None = None) -> int:
    """CUDA device for `local_rank` when a job uses fewer GPUs than the box has.

    On the 8-GPU boxes of this pool the guest shows ONE NUMA node, but the four GPUs on PCI buses below 0x80
    share a ~116 GB/s path to the pinned host memory while the other four do not (profiles/r02_topo_probe.json:
    {0,1,2,3} 116 GB/s aggregate, {4,5,6,7} 209, {0,1,4,5} 213).  A 2- or 4-rank job therefore spreads its ranks
    over both halves instead of taking devices 0..N-1.  With every GPU in use (or no PCI information) the
    mapping is the identity."""
    buses = _pci_bus_numbers() if buses is None else buses
    if not buses or world_size >= len(buses) or local_rank >= world_size:
        return local_rank
    lo = [i for i, b in enumerate(buses) if b < 0x80]
    hi = [i for i, b in enumerate(buses) if b >= 0x80]
    if not lo or not hi:
        return local_rank
    order = []
    for k in range(max(len(lo), len(hi))):          # 0, 4, 1, 5, ... : alternate between the two halves
        if k < len(lo):
            order.append(lo[k])
        if k < len(hi):
            order.append(hi[k])
    return sorted(order[:world_size])[local_rank]
