"""Host-side placement helpers for callers that hand pinned host buffers to the batch searches.

On a two-socket host a pinned buffer that lives on the socket the GPU is NOT attached to crosses the
inter-socket link on every H2D / D2H copy. Linux places pages on the node of the thread that first touches
them, so binding the process to the CPUs next to the GPU before allocating its buffers is enough; nothing
here touches the device. (The GPU boxes of this pool are single-node VMs — profiles/r1/numa_check.txt — so
the helper is a no-op there; the PCIe floor still differs between boxes, 1.67 vs 2.66 ms for 86 MB up +
58 MB down, which is why bench.py reports it next to the end-to-end time.)
"""
import os


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device=0):
    """NUMA node of CUDA device `device` from sysfs, or None when it cannot be told (no sysfs entry,
    single-node box reporting -1)."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device)
        pci = "%04x:%02x:%02x.0" % (bus.pci_domain_id, bus.pci_bus_id, bus.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % pci).read())
        return node if node >= 0 else None
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


def bind_to_gpu_node(device=0):
    """Restricts the calling process to the CPUs of the GPU's NUMA node (intersected with the CPUs it may
    already use; untouched if that would leave none). Returns a dict describing what happened."""
    info = {"numa_node": gpu_numa_node(device), "bound": False}
    if info["numa_node"] is None or not hasattr(os, "sched_setaffinity"):
        return info
    try:
        local = _parse_cpulist(open("/sys/devices/system/node/node%d/cpulist" % info["numa_node"]).read())
        allowed = os.sched_getaffinity(0)
        keep = allowed & local
        info["cpus_before"], info["cpus_local"] = len(allowed), len(keep)
        if keep and keep != allowed:
            os.sched_setaffinity(0, keep)
            info["bound"] = True
    except (OSError, ValueError):
        pass
    return info
