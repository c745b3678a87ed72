"""Topology scheduler daemon: polls gated Pending pods, places each job's gang, pins pods with required nodeAffinity
and removes the gate (reference: gke-topology-scheduler/schedule-daemon.py:447-497,568-810).

    python -m container_engine_accelerators_b200.scheduler.daemon --gate gke.io/topology-aware-auto- --interval 1.0
Fixes: `--ignored-namespace` is honoured (parsed but unused in the reference, schedule-daemon.py:765-767); NotReady nodes are
skipped, not a loop `break` (legacy copy :162-165); cool-offs are parameters so tests run instantly.
"""
from __future__ import annotations

import argparse
import logging
import sys
import time

from ..agent.kube import KubeClient, KubeError
from . import topology as topo

log = logging.getLogger("b200-topology-scheduler")

DEFAULT_GATE_PREFIX = "gke.io/topology-aware-auto-"


def schedule_pod_on_node(kube: KubeClient, pod_name: str, namespace: str, node: dict, gate_name: str) -> bool:
    try:
        pod = kube.get_pod(namespace, pod_name)
        spec = pod.setdefault("spec", {})
        gates = spec.get("schedulingGates") or []
        if any(g.get("name") == gate_name for g in gates):
            spec["affinity"] = {"nodeAffinity": {"requiredDuringSchedulingIgnoredDuringExecution": {"nodeSelectorTerms": [
                {"matchExpressions": [{"key": topo.HOSTNAME_LABEL, "operator": "In", "values": [node["name"]]}]}]}}}
            spec["schedulingGates"] = [g for g in gates if g.get("name") != gate_name]
            kube.replace_pod(namespace, pod_name, pod)
            log.info("Pod %s/%s scheduled on %s with topology %s", namespace, pod_name, node["name"], topo.node_topology_key(node))
    except KubeError as e:
        log.error("Exception when removing pod %s scheduling gate: %s", pod_name, e)
        return False
    return True


def schedule_pods_with_gate(kube: KubeClient, gate_name: str, ignored_namespaces=(), legacy_key: bool = False) -> dict:
    """Returns {job: [(pod, node)]} for what was placed."""
    pending = [p for p in kube.list_pods("status.phase=Pending") if (p["metadata"].get("namespace") or "default") not in ignored_namespaces]
    gated = [p for p in pending if gate_name in {g.get("name") for g in (p.get("spec") or {}).get("schedulingGates") or []}]
    groups = topo.group_pods_by_job(gated)
    log.info("Start scheduling %d jobs", len(groups))
    nodes = kube.list_nodes()
    running = kube.list_pods("status.phase=Running")
    placed: dict = {}
    taken_nodes: set = set()
    for job in sorted(groups, key=lambda j: groups[j][0]["metadata"].get("creationTimestamp") or ""):      # oldest job first
        pods = groups[job]
        try:
            infos = {p["metadata"]["name"]: topo.pod_info(p, job) for p in pods}
            avail = [n for n in nodes if n["metadata"]["name"] not in taken_nodes]
            node_infos = topo.find_schedulable_nodes(avail, running, next(iter(infos.values()))["tolerations"])
            if len(infos) > len(node_infos):
                log.error("Not enough nodes available for job %s scheduling: %d nodes, %d pods. Skipping job.", job, len(node_infos), len(infos))
                continue
            sorted_pods = sorted(infos.values(), key=topo.pod_sorting_key)
            sorted_nodes = sorted(node_infos.values(), key=lambda n: topo.node_topology_key(n, legacy_key))
            assignment = topo.calculate_pods_assignment(sorted_nodes, sorted_pods, legacy_key)
            if not assignment:
                log.error("No scheduling for job %s with gate %s was found. Skipping job.", job, gate_name)
                continue
            for i, pod in enumerate(sorted_pods):
                node = sorted_nodes[assignment[i]]
                if not schedule_pod_on_node(kube, pod["name"], pod["namespace"], node, gate_name):
                    log.error("Failed to schedule pod %s on node %s. Skipping job %s", pod["name"], node["name"], job)
                    break
                taken_nodes.add(node["name"])
                placed.setdefault(job, []).append((pod["name"], node["name"]))
        except Exception as e:     # one bad job must not stop the others
            log.exception("Exception when scheduling %s job, gate %s: %s", job, gate_name, e)
    return placed


def run_scheduling_loop(kube: KubeClient, gate_prefix: str = DEFAULT_GATE_PREFIX, interval: float = 1.0, ignored_namespaces=(), legacy_key: bool = False,
                        startup_cooloff: float = 90.0, gang_settle: float = 5.0, gate_cooloff: float = 60.0, iterations=None, sleep=time.sleep) -> None:
    log.info("[Cool off] %ssec", startup_cooloff)
    sleep(startup_cooloff)     # after a restart, let previously placed pods show up on nodes before estimating free resources
    last = time.time() - interval
    n = 0
    while iterations is None or n < iterations:
        n += 1
        wait = interval - (time.time() - last)
        if wait > 0:
            sleep(wait)
        last = time.time()
        try:
            pods = [p for p in kube.list_pods("status.phase=Pending") if (p["metadata"].get("namespace") or "default") not in ignored_namespaces]
            gates = topo.find_pod_gates(pods, gate_prefix)
            log.info("Found %d pending pods and %d gates", len(pods), len(gates))
            if not gates:
                continue
            sleep(gang_settle)             # let the whole gang appear
            for g in sorted(gates):
                log.info("Scheduling pods with gate %s", g)
                schedule_pods_with_gate(kube, g, ignored_namespaces, legacy_key)
                sleep(gate_cooloff)
        except KubeError as e:
            log.error("Exception when listing Kubernetes nodes or pods: %s", e)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="b200-topology-scheduler")
    ap.add_argument("-g", "--gate", default=DEFAULT_GATE_PREFIX)
    ap.add_argument("-i", "--interval", type=float, default=1.0)
    ap.add_argument("--ignored-namespace", nargs="*", default=[])
    ap.add_argument("--legacy-placement-group-key", action="store_true", help="4-level key with gke-placement-group first (older TCPXO variant)")
    ap.add_argument("--kube-url", default="", help="API server URL (default: B200_KUBE_URL, else in-cluster)")
    ap.add_argument("--startup-cooloff", type=float, default=90.0, help="seconds to wait after start so pods placed before a restart show up on their nodes")
    ap.add_argument("--gang-settle", type=float, default=5.0, help="seconds to let the rest of a gang appear once a gate is seen")
    ap.add_argument("--gate-cooloff", type=float, default=60.0, help="seconds to wait after scheduling a gate")
    ap.add_argument("--iterations", type=int, default=None, help=argparse.SUPPRESS)
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(levelname).1s %(name)s] %(message)s")
    kube = KubeClient.from_env(args.kube_url)
    run_scheduling_loop(kube, args.gate, args.interval, tuple(args.ignored_namespace), args.legacy_placement_group_key, startup_cooloff=args.startup_cooloff,
                        gang_settle=args.gang_settle, gate_cooloff=args.gate_cooloff, iterations=args.iterations)
    return 0


if __name__ == "__main__":
    sys.exit(main())
