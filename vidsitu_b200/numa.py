"""Host-side placement for the pinned staging buffers of the end-to-end path (one process per
GPU, like utils/trn_dist_utils.py:5-42).  cudaHostAlloc places pages on the NUMA node of the allocating thread; with
eight ranks copying 308 MB per step each, where those pages live decides whether the H2D copies share one memory
controller.  Everything here degrades to a no-op when the topology cannot be read."""
from __future__ import annotations

import contextlib
import glob
import os
import re
from typing import Dict, List, Optional


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def numa_nodes() -> Dict[int, List[int]]:
    """NUMA node -> CPUs of that node this process may run on."""
    try:
        allowed = set(os.sched_getaffinity(0))
    except AttributeError:
        return {}
    nodes: Dict[int, List[int]] = {}
    for path in sorted(glob.glob("/sys/devices/system/node/node[0-9]*/cpulist")):
        m = re.search(r"node(\d+)/cpulist$", path)
        try:
            cpus = [c for c in _parse_cpulist(open(path).read()) if c in allowed]
        except OSError:
            continue
        if cpus:
            nodes[int(m.group(1))] = cpus
    return nodes


def gpu_cpu_affinity(gpu_index: int) -> Optional[List[int]]:
    """CPUs the driver reports as local to the GPU (nvidia-smi topo 'CPU Affinity')."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, v in enumerate(mask) for b in range(64) if (int(v) >> b) & 1]
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        return cpus or None
    except Exception:  # noqa: BLE001 -- no NVML / not permitted: leave the process where it is
        return None


def bind_to_gpu(gpu_index: int) -> Optional[List[int]]:
    """Pin the calling process to the CPUs local to `gpu_index`; returns the CPU list or None."""
    cpus = gpu_cpu_affinity(gpu_index)
    if cpus:
        try:
            os.sched_setaffinity(0, cpus)
        except OSError:
            return None
    return cpus


@contextlib.contextmanager
def on_node(node: Optional[int]):
    """Run the body with the thread restricted to one NUMA node's CPUs (first-touch placement of what it allocates)."""
    nodes = numa_nodes()
    if node is None or node not in nodes:
        yield False
        return
    old = os.sched_getaffinity(0)
    try:
        os.sched_setaffinity(0, nodes[node])
        yield True
    finally:
        os.sched_setaffinity(0, old)


def node_of_cpus(cpus: List[int]) -> Optional[int]:
    for n, cs in numa_nodes().items():
        if cpus and cpus[0] in cs:
            return n
    return None
