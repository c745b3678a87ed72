"""GetPreferredAllocation: which of the free devices the kubelet should hand to a container.

The reference registers without options and leaves the choice to the kubelet, which picks arbitrary ids
(pkg/gpu/nvidia/beta_plugin.go:35-37,100-103). That is harmless for whole GPUs behind one NVSwitch (every pair is
equidistant) but not for
  * shared GPUs: "any free vGPU id" can stack three pods on one physical GPU while the others idle;
  * multi-GPU requests on a two-socket host: GPUs from one NUMA node keep host staging buffers and NIC traffic local.
Opt-in (`--preferred-allocation-policy spread|packed`; default `none` keeps the reference's contract bit for bit).

Selection, one device at a time:
  1. ids the kubelet says must be included come first;
  2. NUMA: stay on the nodes already used; when a new node is needed take the smallest one that covers what is still
     missing, else the one with most free devices;
  3. within a node, sharing policy: `spread` takes the physical GPU with the most free replicas and avoids physical GPUs
     already chosen for this container; `packed` takes the one with the fewest free replicas (leaves whole GPUs free);
  4. ties: natural id order (nvidia2 before nvidia10).
"""
from __future__ import annotations

import re
from typing import Callable, Optional

from . import sharing

POLICIES = ("none", "spread", "packed")


def natural_key(device_id: str) -> list:
    return [int(tok) if tok.isdigit() else tok for tok in re.split(r"(\d+)", device_id)]


def physical_of(device_id: str) -> str:
    return sharing.virtual_to_physical_device_id(device_id) if sharing.is_virtual_device_id(device_id) else device_id


def preferred_allocation(available: list, must_include: list, size: int, numa_of: Callable[[str], Optional[int]], policy: str = "spread") -> list:
    """Returns `size` device ids (fewer only if fewer exist). Pure function: the service passes the kubelet's lists."""
    if policy not in POLICIES or policy == "none":
        raise ValueError(f"unknown preferred-allocation policy {policy!r}")
    chosen: list = []
    for d in must_include:
        if d not in chosen:
            chosen.append(d)
    pool = sorted({d for d in available if d not in chosen}, key=natural_key)
    while len(chosen) < size and pool:
        missing = size - len(chosen)
        used_nodes = {numa_of(d) for d in chosen}
        free_per_node: dict = {}
        for d in pool:
            free_per_node[numa_of(d)] = free_per_node.get(numa_of(d), 0) + 1
        free_per_phys: dict = {}
        for d in pool:
            free_per_phys[physical_of(d)] = free_per_phys.get(physical_of(d), 0) + 1
        chosen_phys = {physical_of(d) for d in chosen}

        def node_rank(node) -> tuple:
            if node in used_nodes:
                return (0, 0)
            free = free_per_node[node]
            return (1, free) if free >= missing else (2, -free)      # smallest node that fits, else the biggest one

        def key(d: str) -> tuple:
            phys = physical_of(d)
            if policy == "spread":
                share = (1 if phys in chosen_phys else 0, -free_per_phys[phys])
            else:
                share = (0 if phys in chosen_phys else 1, free_per_phys[phys])
            return (node_rank(numa_of(d)), share, natural_key(d))

        best = min(pool, key=key)
        chosen.append(best)
        pool.remove(best)
    return chosen[:max(size, len(must_include))]
