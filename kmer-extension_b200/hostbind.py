"""Host-side placement for one-process-per-GPU hosts: run on the CPUs of the NUMA node the GPU hangs off.

Pinned host buffers (cudaHostAlloc, torch's pin_memory) are placed on the node of the thread that allocates them.  A rank
that runs on the other socket puts its buffers there, and every host<->device copy then crosses the socket interconnect:
with eight GPUs copying results back at once that link, not PCIe, sets the pace.  Call bind_to_gpu() before the first
pinned allocation.  (The reference has no counterpart: PostgreSQL backends are placed by the OS.)"""
from __future__ import annotations

import os


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_sysfs_dir(device_index: int) -> str | None:
    import torch
    p = torch.cuda.get_device_properties(device_index)
    try:
        name = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except AttributeError:
        return None
    d = f"/sys/bus/pci/devices/{name}"
    return d if os.path.isdir(d) else None


def bind_to_gpu(device_index: int) -> dict:
    """Restricts this process (the calling thread and the threads it starts later) to the CPUs local to the GPU.
    Returns what was done: {"numa_node", "cpus", "bound", "why"}; never raises -- placement is an optimisation."""
    info = {"numa_node": None, "cpus": None, "bound": False, "why": ""}
    try:
        d = gpu_sysfs_dir(device_index)
        if d is None:
            info["why"] = "no sysfs entry for the GPU"
            return info
        with open(os.path.join(d, "numa_node")) as f:
            info["numa_node"] = int(f.read().strip())
        with open(os.path.join(d, "local_cpulist")) as f:
            local = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        target = local & allowed
        if not target:
            info["why"] = "none of the GPU's local CPUs is in this process's cpuset"
            return info
        info["cpus"] = len(target)
        if target == allowed:
            info["why"] = "already local (single node, or placed by the launcher)"
            info["bound"] = True
            return info
        os.sched_setaffinity(0, target)
        info["bound"] = True
    except Exception as e:  # noqa: BLE001
        info["why"] = f"{type(e).__name__}: {e}"
    return info
