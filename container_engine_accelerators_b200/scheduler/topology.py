"""Topology-aware gang placement: pure functions (no API calls) so every rule is unit-testable.

Behaviour follows the reference scheduler (gke-topology-scheduler/schedule-daemon.py):
label precedence GA `cloud.google.com/gce-topology-{block,subblock,host}` then `topology.gke.io/{cluster,rack,host}`;
distance 1e6 / 1e4 / 1e2 at the first differing level, 0 when equal or unlabeled; cost = sum over consecutive pods in
index order; pods ordered by completion index -> kubeflow replica index -> numeric name suffix; nodes filtered by taints,
readiness and free cpu/memory/gpu; each pod lands on a distinct node, node order preserved.

Differences: the assignment is an exact O(pods x nodes^2) dynamic programme instead of the reference's odometer over all
C(nodes, pods) combinations (schedule-daemon.py:500-544) — same optimum, polynomial time; pods and nodes are plain dicts
(REST JSON) because the image has no `kubernetes` package. `legacy_key=True` reproduces the older TCPXO variant's 4-level
key with `cloud.google.com/gke-placement-group` first (gpudirect-tcpxo/topology-scheduler/schedule-daemon.py:74-91).
"""
from __future__ import annotations

import logging
import re

from .quantity import parse_quantity

log = logging.getLogger("b200-topology-scheduler")

PRERELEASE_LABELS = ("topology.gke.io/cluster", "topology.gke.io/rack", "topology.gke.io/host")
GA_LABELS = ("cloud.google.com/gce-topology-block", "cloud.google.com/gce-topology-subblock", "cloud.google.com/gce-topology-host")
PLACEMENT_GROUP_LABEL = "cloud.google.com/gke-placement-group"
JOB_COMPLETION_INDEX_LABEL = "batch.kubernetes.io/job-completion-index"
JOB_NAME_LABEL = "job-name"
KUBEFLOW_REPLICA_INDEX_LABEL = "training.kubeflow.org/replica-index"
KUBEFLOW_JOB_NAME_LABEL = "training.kubeflow.org/job-name"
UNKNOWN_JOB = "jobless"
GPU_RESOURCE = "nvidia.com/gpu"
HOSTNAME_LABEL = "kubernetes.io/hostname"


# ------------------------------------------------------------------------------------------------- job identity
def _labels(pod: dict) -> dict:
    return (pod.get("metadata") or {}).get("labels") or {}


def job_name_from_label(pod: dict) -> str:
    return _labels(pod).get(JOB_NAME_LABEL, UNKNOWN_JOB)


def job_name_from_kubeflow(pod: dict) -> str:
    return _labels(pod).get(KUBEFLOW_JOB_NAME_LABEL, UNKNOWN_JOB)


def job_name_from_owner(pod: dict) -> str:
    refs = (pod.get("metadata") or {}).get("ownerReferences") or []
    return refs[0].get("uid", UNKNOWN_JOB) if refs else UNKNOWN_JOB


def job_name_from_helm(pod: dict) -> str:
    return _labels(pod).get("name", UNKNOWN_JOB)


JOB_NAME_EXTRACTORS = (job_name_from_label, job_name_from_kubeflow, job_name_from_owner, job_name_from_helm)


# ------------------------------------------------------------------------------------------------- keys & distance
def node_topology_key(node_info: dict, legacy_key: bool = False) -> tuple:
    labels = node_info.get("node_labels") or {}
    for trio in (GA_LABELS, PRERELEASE_LABELS):
        if all(l in labels for l in trio):
            key = tuple(labels[l] for l in trio)
            if legacy_key:
                return (labels.get(PLACEMENT_GROUP_LABEL, ""),) + key
            return key
    return ()


def node_topology_distance(a: dict, b: dict, legacy_key: bool = False) -> float:
    ka, kb = node_topology_key(a, legacy_key), node_topology_key(b, legacy_key)
    result = 1000000.0 * (100.0 if legacy_key else 1.0)
    for i, part in enumerate(ka):
        if i >= len(kb) or part != kb[i]:
            return result
        result /= 100.0
    return 0.0


def pod_sorting_key(pod_info: dict):
    """(0, index) for indexed pods, (1, prefix, number) for name-ordered ones — homogeneous within a job in practice."""
    idx = pod_info.get("index")
    if idx is not None:
        try:
            return (0, "", int(idx))
        except (ValueError, TypeError):
            log.error("Error converting %s pod index %r to integer", pod_info.get("name"), idx)
    name = pod_info["name"]
    m = re.fullmatch(r"(.*?)(\d+)", name)
    if m:
        return (1, m.group(1), int(m.group(2)))
    log.warning("Pod %s does not have a numeric suffix. Using 0 as index.", name)
    return (1, name, 0)


# ------------------------------------------------------------------------------------------------- resources
def _container_requests(container: dict):
    req = ((container.get("resources") or {}).get("requests")) or {}
    return parse_quantity(req.get("cpu", 0)), parse_quantity(req.get("memory", 0)), int(req.get(GPU_RESOURCE, 0))


def pod_used_resources(pod: dict):
    """Requests of containers that are not terminated (a Running pod's live footprint)."""
    statuses = (pod.get("status") or {}).get("containerStatuses")
    if not statuses:
        return 0, 0, 0
    cpu = mem = gpu = 0
    for c, st in zip((pod.get("spec") or {}).get("containers") or [], statuses):
        if ((st.get("state") or {}).get("terminated")) is not None:
            continue
        c_cpu, c_mem, c_gpu = _container_requests(c)
        cpu += c_cpu; mem += c_mem; gpu += c_gpu
    return cpu, mem, gpu


def pod_info(pod: dict, job_name: str) -> dict:
    meta, spec = pod.get("metadata") or {}, pod.get("spec") or {}
    labels = meta.get("labels") or {}
    index = labels.get(JOB_COMPLETION_INDEX_LABEL, labels.get(KUBEFLOW_REPLICA_INDEX_LABEL))
    cpu = mem = gpu = 0
    for c in spec.get("containers") or []:
        c_cpu, c_mem, c_gpu = _container_requests(c)
        cpu += c_cpu; mem += c_mem; gpu += c_gpu
    info = {"name": meta.get("name"), "namespace": meta.get("namespace", "default"), "index": index, "cpu": cpu, "memory": mem, "gpu": gpu,
            "tolerations": spec.get("tolerations") or [], "job_name": job_name, "creation_time": meta.get("creationTimestamp")}
    if spec.get("nodeSelector"):
        info["node_selector"] = spec["nodeSelector"]
    return info


def tolerates(taints: list, tolerations: list) -> bool:
    """The reference's rule: every taint key must appear in the tolerations; `Equal` must also match the value."""
    by_key = {t.get("key"): t for t in tolerations or []}
    for taint in taints or []:
        tol = by_key.get(taint.get("key"))
        if tol is None:
            return False
        if tol.get("operator") == "Equal" and tol.get("value") != taint.get("value"):
            return False
    return True


def node_is_ready(node: dict) -> bool:
    """NotReady nodes are skipped (the legacy variant `break`s out of the whole loop here — SURVEY §7.3-9 bug, not copied)."""
    for c in (node.get("status") or {}).get("conditions") or []:
        if c.get("type") == "Ready" and c.get("status") != "True":
            return False
    return True


def find_schedulable_nodes(nodes: list, running_pods: list, tolerations: list) -> dict:
    used: dict = {}
    for p in running_pods:
        n = (p.get("spec") or {}).get("nodeName")
        if n:
            cpu, mem, gpu = pod_used_resources(p)
            u = used.setdefault(n, [0, 0, 0])
            u[0] += cpu; u[1] += mem; u[2] += gpu
    out = {}
    for node in nodes:
        name = node["metadata"]["name"]
        if not tolerates((node.get("spec") or {}).get("taints") or [], tolerations):
            log.info("Skipping node %s because it has taints not covered by pod tolerations", name)
            continue
        if not node_is_ready(node):
            log.info("Skipping node %s because it is NotReady", name)
            continue
        alloc = (node.get("status") or {}).get("allocatable") or {}
        u = used.get(name, [0, 0, 0])
        info = {"name": name, "cpu": parse_quantity(alloc.get("cpu", 0)) - u[0], "memory": parse_quantity(alloc.get("memory", 0)) - u[1],
                "gpu": int(alloc.get(GPU_RESOURCE, 0)) - u[2]}
        if node["metadata"].get("labels"):
            info["node_labels"] = node["metadata"]["labels"]
        out[name] = info
    return out


def can_schedule(node: dict, pod: dict) -> bool:
    labels = node.get("node_labels") or {}
    for k, v in (pod.get("node_selector") or {}).items():
        if labels.get(k) != v:
            return False
    return node["cpu"] >= pod["cpu"] and node["memory"] >= pod["memory"] and node["gpu"] >= pod["gpu"]


# ------------------------------------------------------------------------------------------------- assignment
def calculate_pods_assignment(sorted_nodes: list, sorted_pods: list, legacy_key: bool = False) -> list:
    """Strictly increasing node indices a[0] < a[1] < ... (one pod per node, node order preserved) minimising
    sum_i dist(node[a[i]], node[a[i-1]]); [] when no feasible assignment exists.
    dp[i][j] = best cost with pod i on node j; ties resolve to the lowest node indices (the first optimum the
    reference's enumeration would meet)."""
    k, n = len(sorted_pods), len(sorted_nodes)
    if k == 0 or k > n:
        return []
    INF = float("inf")
    feasible = [[can_schedule(sorted_nodes[j], sorted_pods[i]) for j in range(n)] for i in range(k)]
    keys = [node_topology_key(nd, legacy_key) for nd in sorted_nodes]

    def dist(a: int, b: int) -> float:
        ka, kb = keys[a], keys[b]
        result = 1000000.0 * (100.0 if legacy_key else 1.0)
        for lvl, part in enumerate(ka):
            if lvl >= len(kb) or part != kb[lvl]:
                return result
            result /= 100.0
        return 0.0
    cost = [[INF] * n for _ in range(k)]
    back = [[-1] * n for _ in range(k)]
    for j in range(n - k + 1):
        if feasible[0][j]:
            cost[0][j] = 0.0
    for i in range(1, k):
        for j in range(i, n - (k - 1 - i)):
            if not feasible[i][j]:
                continue
            best, arg = INF, -1
            for jp in range(i - 1, j):
                c = cost[i - 1][jp]
                if c == INF:
                    continue
                c += dist(j, jp)
                if c < best:
                    best, arg = c, jp
            cost[i][j], back[i][j] = best, arg
    end, best = -1, INF
    for j in range(k - 1, n):
        if cost[k - 1][j] < best:
            best, end = cost[k - 1][j], j
    if end < 0:
        return []
    out = [0] * k
    j = end
    for i in range(k - 1, -1, -1):
        out[i] = j
        j = back[i][j]
    return out


def assignment_cost(sorted_nodes: list, assignment: list, legacy_key: bool = False) -> float:
    return sum(node_topology_distance(sorted_nodes[assignment[i]], sorted_nodes[assignment[i - 1]], legacy_key) for i in range(1, len(assignment)))


# ------------------------------------------------------------------------------------------------- grouping
def find_pod_gates(pods: list, prefix: str) -> set:
    out = set()
    for p in pods:
        for g in (p.get("spec") or {}).get("schedulingGates") or []:
            if g.get("name", "").startswith(prefix):
                out.add(g["name"])
    return out


def group_pods_by_job(pods: list) -> dict:
    """Extractors tried in order; a pod belongs to the first group that names it; leftovers are scheduled together."""
    groups: dict = {}
    unassigned = {p["metadata"]["name"] for p in pods}
    for extract in JOB_NAME_EXTRACTORS:
        buckets: dict = {}
        for p in pods:
            buckets.setdefault(extract(p), []).append(p)
        for job, members in buckets.items():
            if job == UNKNOWN_JOB:
                continue
            names = {p["metadata"]["name"] for p in members}
            if not all(p["metadata"].get("creationTimestamp") for p in members):
                log.error("No pod creationTimestamp in job %s. Job ignored.", job)
                unassigned -= names
                continue
            tol0 = (members[0].get("spec") or {}).get("tolerations")
            if not all(((p.get("spec") or {}).get("tolerations")) == tol0 for p in members[1:]):
                log.error("Pods in job %s have different tolerations. Job ignored.", job)
                unassigned -= names
                continue
            if not (names & unassigned):
                continue        # every pod already claimed by an earlier extractor
            unassigned -= names
            groups[job] = members
    if unassigned:
        log.warning("Found %d pods without explicit job name, going to schedule all together.", len(unassigned))
        groups["pods-without-explicit-job-name"] = [p for p in pods if p["metadata"]["name"] in unassigned]
    return groups
