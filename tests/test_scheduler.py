"""Topology scheduler + labeler. The reference ships zero tests for these; rules come from Appendix A.8."""
import itertools

import pytest

from container_engine_accelerators_b200.agent import kube, testing
from container_engine_accelerators_b200.scheduler import daemon, labeler
from container_engine_accelerators_b200.scheduler import topology as topo
from container_engine_accelerators_b200.scheduler.quantity import parse_quantity


def node(name, block=None, sub=None, host=None, gpu=8, cpu="200", mem="1000Gi", ready=True, taints=None, labels=None, prerelease=False):
    lab = dict(labels or {})
    if block is not None:
        keys = topo.PRERELEASE_LABELS if prerelease else topo.GA_LABELS
        lab.update(dict(zip(keys, (block, sub, host))))
    return {"metadata": {"name": name, "labels": lab}, "spec": {"taints": taints or []},
            "status": {"conditions": [{"type": "Ready", "status": "True" if ready else "False"}], "allocatable": {"cpu": cpu, "memory": mem, "nvidia.com/gpu": str(gpu)}}}


def pod(name, job="j1", index=None, gpu=8, gate="gke.io/topology-aware-auto-j1", ns="default", phase="Pending", node_name=None, tolerations=None, ts="2026-01-01T00:00:00Z", labels=None):
    lab = {"job-name": job} if job else {}
    if index is not None:
        lab[topo.JOB_COMPLETION_INDEX_LABEL] = str(index)
    lab.update(labels or {})
    p = {"metadata": {"name": name, "namespace": ns, "labels": lab, "creationTimestamp": ts},
         "spec": {"containers": [{"name": "c", "resources": {"requests": {"cpu": "10", "memory": "100Gi", "nvidia.com/gpu": str(gpu)}}}], "tolerations": tolerations or []},
         "status": {"phase": phase}}
    if gate:
        p["spec"]["schedulingGates"] = [{"name": gate}]
    if node_name:
        p["spec"]["nodeName"] = node_name
        p["status"]["containerStatuses"] = [{"state": {"running": {}}}]
    return p


def test_parse_quantity():
    assert parse_quantity("100m") == parse_quantity("0.1") and parse_quantity("2Gi") == 2 * 1024 ** 3 and parse_quantity("1500M") == 1500 * 10 ** 6 and parse_quantity(3) == 3
    with pytest.raises(ValueError):
        parse_quantity("abc")


def test_label_precedence_and_distance():
    a = {"name": "a", "node_labels": {**dict(zip(topo.GA_LABELS, "b1 s1 h1".split())), **dict(zip(topo.PRERELEASE_LABELS, "x y z".split()))}}
    assert topo.node_topology_key(a) == ("b1", "s1", "h1")                         # GA labels win
    mk = lambda b, s, h: {"name": "n", "node_labels": dict(zip(topo.GA_LABELS, (b, s, h)))}
    assert topo.node_topology_distance(mk("b1", "s1", "h1"), mk("b2", "s1", "h1")) == 1e6
    assert topo.node_topology_distance(mk("b1", "s1", "h1"), mk("b1", "s2", "h1")) == 1e4
    assert topo.node_topology_distance(mk("b1", "s1", "h1"), mk("b1", "s1", "h2")) == 1e2
    assert topo.node_topology_distance(mk("b1", "s1", "h1"), mk("b1", "s1", "h1")) == 0
    assert topo.node_topology_distance({"name": "u"}, mk("b1", "s1", "h1")) == 0    # unlabeled => key () => distance 0
    assert topo.node_topology_key({"name": "p", "node_labels": dict(zip(topo.PRERELEASE_LABELS, "c r h".split()))}) == ("c", "r", "h")


def test_pod_ordering():
    infos = [{"name": "w-10", "index": None}, {"name": "w-2", "index": None}, {"name": "w-1", "index": None}]
    assert [p["name"] for p in sorted(infos, key=topo.pod_sorting_key)] == ["w-1", "w-2", "w-10"]
    infos = [{"name": "a", "index": "3"}, {"name": "b", "index": "0"}, {"name": "c", "index": "12"}]
    assert [p["name"] for p in sorted(infos, key=topo.pod_sorting_key)] == ["b", "a", "c"]
    p = topo.pod_info(pod("x", labels={topo.KUBEFLOW_REPLICA_INDEX_LABEL: "4"}), "j")
    assert p["index"] == "4" and p["gpu"] == 8


def test_node_filters():
    nodes = [node("ok", "b", "s", "h1"), node("notready", "b", "s", "h2", ready=False), node("tainted", "b", "s", "h3", taints=[{"key": "k", "value": "v", "effect": "NoSchedule"}]),
             node("gpu-taint", "b", "s", "h4", taints=[{"key": "nvidia.com/gpu", "value": "present", "effect": "NoSchedule"}]), node("busy", "b", "s", "h5")]
    running = [pod("r", gate=None, phase="Running", node_name="busy", gpu=6)]
    tol = [{"key": "nvidia.com/gpu", "operator": "Exists"}]
    got = topo.find_schedulable_nodes(nodes, running, tol)
    assert set(got) == {"ok", "gpu-taint", "busy"} and got["busy"]["gpu"] == 2 and got["ok"]["gpu"] == 8
    assert not topo.tolerates([{"key": "k", "value": "v"}], [{"key": "k", "operator": "Equal", "value": "other"}])
    assert topo.tolerates([{"key": "k", "value": "v"}], [{"key": "k", "operator": "Equal", "value": "v"}])
    assert not topo.can_schedule(got["busy"], {"cpu": 1, "memory": 1, "gpu": 8})
    assert not topo.can_schedule({**got["ok"], "node_labels": {"a": "b"}}, {"cpu": 1, "memory": 1, "gpu": 1, "node_selector": {"a": "c"}})


def brute_force(sorted_nodes, sorted_pods):
    best, best_cost = [], float("inf")
    for combo in itertools.combinations(range(len(sorted_nodes)), len(sorted_pods)):
        if all(topo.can_schedule(sorted_nodes[j], sorted_pods[i]) for i, j in enumerate(combo)):
            c = topo.assignment_cost(sorted_nodes, list(combo))
            if c < best_cost:
                best, best_cost = list(combo), c
    return best, best_cost


def test_assignment_matches_exhaustive_search():
    import random
    rng = random.Random(7)
    for trial in range(40):
        n, k = rng.randint(3, 9), rng.randint(1, 4)
        nodes = [{"name": f"n{i}", "cpu": 100, "memory": 100, "gpu": rng.choice([0, 4, 8]),
                  "node_labels": dict(zip(topo.GA_LABELS, (f"b{rng.randint(0, 1)}", f"s{rng.randint(0, 2)}", f"h{i}")))} for i in range(n)]
        nodes.sort(key=topo.node_topology_key)
        pods = [{"name": f"p{i}", "cpu": 1, "memory": 1, "gpu": rng.choice([4, 8])} for i in range(k)]
        got = topo.calculate_pods_assignment(nodes, pods)
        want, want_cost = brute_force(nodes, pods)
        if not want:
            assert got == []
        else:
            assert got == sorted(got) and len(set(got)) == k
            assert topo.assignment_cost(nodes, got) == want_cost, (trial, got, want)


def test_assignment_prefers_compact_placement():
    nodes = sorted([{"name": f"n{i}", "cpu": 10, "memory": 10, "gpu": 8, "node_labels": dict(zip(topo.GA_LABELS, lab))} for i, lab in enumerate(
        [("b1", "s1", "h1"), ("b2", "s1", "h1"), ("b2", "s1", "h2"), ("b2", "s2", "h1"), ("b3", "s1", "h1")])], key=topo.node_topology_key)
    pods = [{"name": f"p{i}", "cpu": 1, "memory": 1, "gpu": 8} for i in range(3)]
    got = topo.calculate_pods_assignment(nodes, pods)
    assert [nodes[j]["node_labels"][topo.GA_LABELS[0]] for j in got] == ["b2", "b2", "b2"]
    assert topo.calculate_pods_assignment(nodes[:2], pods) == []


def test_job_grouping_order_and_leftovers():
    pods = [pod("a-0", job="ja"), pod("a-1", job="ja"), pod("k-0", job=None, labels={topo.KUBEFLOW_JOB_NAME_LABEL: "kf"}),
            {**pod("o-0", job=None), "metadata": {**pod("o-0", job=None)["metadata"], "ownerReferences": [{"uid": "uid-1"}]}}, pod("h-0", job=None, labels={"name": "helm"}), pod("lonely", job=None)]
    groups = topo.group_pods_by_job(pods)
    assert {k: [p["metadata"]["name"] for p in v] for k, v in groups.items()} == {"ja": ["a-0", "a-1"], "kf": ["k-0"], "uid-1": ["o-0"], "helm": ["h-0"], "pods-without-explicit-job-name": ["lonely"]}
    mixed = [pod("m-0", job="jm"), pod("m-1", job="jm", tolerations=[{"key": "x", "operator": "Exists"}])]
    assert topo.group_pods_by_job(mixed) == {}                                      # different tolerations: job ignored


def _scheduler_role():
    import os
    import yaml
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return next(d for d in yaml.safe_load_all(open(os.path.join(root, "deploy", "topology-scheduler", "service-account.yaml"))) if d["kind"] == "ClusterRole")


@pytest.fixture
def api():
    a = testing.FakeKubeApi().start()
    yield a
    a.stop()


def test_end_to_end_bind_sets_affinity_and_removes_gate(api):
    for i, lab in enumerate([("b1", "s1", "h1"), ("b1", "s1", "h2"), ("b2", "s1", "h1")]):
        n = node(f"node{i}", *lab)
        api.nodes[n["metadata"]["name"]] = n
    gate = "gke.io/topology-aware-auto-train"
    for i in range(2):
        api.add_pod(pod(f"train-{i}", job="train", index=i, gate=gate))
    api.add_pod(pod("ignored-0", job="other", index=0, gate=gate, ns="kube-system"))
    api.add_pod(pod("busy", gate=None, phase="Running", node_name="node2", gpu=8))
    kc = kube.KubeClient(api.url)
    assert topo.find_pod_gates(kc.list_pods("status.phase=Pending"), daemon.DEFAULT_GATE_PREFIX) == {gate}
    placed = daemon.schedule_pods_with_gate(kc, gate, ignored_namespaces=("kube-system",))
    assert placed == {"train": [("train-0", "node0"), ("train-1", "node1")]}
    for i in range(2):
        spec = api.pods[("default", f"train-{i}")]["spec"]
        assert spec["schedulingGates"] == []
        term = spec["affinity"]["nodeAffinity"]["requiredDuringSchedulingIgnoredDuringExecution"]["nodeSelectorTerms"][0]["matchExpressions"][0]
        assert term == {"key": "kubernetes.io/hostname", "operator": "In", "values": [f"node{i}"]}
    assert api.pods[("kube-system", "ignored-0")]["spec"]["schedulingGates"] == [{"name": gate}]      # --ignored-namespace is honoured
    assert testing.rbac_violations(_scheduler_role(), api.requests) == []                            # covered by deploy/topology-scheduler/service-account.yaml


def test_not_enough_nodes_skips_job(api):
    api.nodes["only"] = node("only", "b", "s", "h")
    gate = "gke.io/topology-aware-auto-x"
    for i in range(2):
        api.add_pod(pod(f"x-{i}", job="x", index=i, gate=gate))
    assert daemon.schedule_pods_with_gate(kube.KubeClient(api.url), gate) == {}


def test_scheduling_loop_runs_with_injected_sleep(api):
    api.nodes["n0"] = node("n0", "b", "s", "h")
    api.add_pod(pod("solo-0", job="solo", index=0, gate="gke.io/topology-aware-auto-solo"))
    sleeps = []
    daemon.run_scheduling_loop(kube.KubeClient(api.url), iterations=1, sleep=sleeps.append)
    assert sleeps[0] == 90.0 and 5.0 in sleeps and 60.0 in sleeps                      # reference cool-offs (A.8)
    assert api.pods[("default", "solo-0")]["spec"]["schedulingGates"] == []


def test_legacy_placement_group_key():
    n1 = {"name": "a", "node_labels": {**dict(zip(topo.GA_LABELS, "b s h".split())), topo.PLACEMENT_GROUP_LABEL: "pg1"}}
    n2 = {"name": "b", "node_labels": {**dict(zip(topo.GA_LABELS, "b s h".split())), topo.PLACEMENT_GROUP_LABEL: "pg2"}}
    assert topo.node_topology_key(n1, legacy_key=True) == ("pg1", "b", "s", "h")
    assert topo.node_topology_distance(n1, n2, legacy_key=True) == 1e8 and topo.node_topology_distance(n1, n2) == 0


def test_labeler_from_metadata_and_nvml(api, tmp_path):
    assert labeler.labels_from_physical_host("/c1/r2/h3") == {"topology.gke.io/cluster": "c1", "topology.gke.io/rack": "r2", "topology.gke.io/host": "h3"}
    with pytest.raises(ValueError):
        labeler.labels_from_physical_host("bogus")

    class FakeMeta:
        def get(self, url, headers=None, timeout=None):
            assert headers == {"Metadata-Flavor": "Google"}
            text = "gke-node-1" if url.endswith("/name") else "/cl/ra/ho"
            return type("R", (), {"status_code": 200, "text": text})()
    api.add_node("gke-node-1", labels={"keep": "1"})
    labeler.update_node_labels_from_metadata(kube.KubeClient(api.url), session=FakeMeta())
    assert api.nodes["gke-node-1"]["metadata"]["labels"] == {"keep": "1", "topology.gke.io/cluster": "cl", "topology.gke.io/rack": "ra", "topology.gke.io/host": "ho"}
    assert testing.rbac_violations(_scheduler_role(), api.requests) == []
    from container_engine_accelerators_b200.agent import nvml
    dev = testing.make_fake_dev(str(tmp_path), 8)
    pci = testing.make_fake_pci(str(tmp_path), "0000:1b:00.0", 1)
    labs = labeler.labels_from_nvml(nvml.MockNvml(dev, bus_id="00000000:1B:00.0"), pci)
    assert labs["b200.gke.io/gpu-count"] == "8" and labs["b200.gke.io/gpu-model"] == "NVIDIA-B200" and labs["b200.gke.io/numa-nodes"] == "1" and len(labs["b200.gke.io/nvlink-domain"]) == 12


def test_expert_dispatch_plan_is_consistent():
    from container_engine_accelerators_b200.models.workloads import expert_dispatch_plan, uniform_plan
    p = expert_dispatch_plan(8, 4096, 1024, top_k=2, skew=1.0, seed=3)
    assert p.rows.shape == (8, 8) and (p.rows.sum(axis=1) == 4096 * 2).all()            # every token goes to top_k experts
    for s in range(8):
        assert list(p.send_off[s]) == [int(p.rows[s, :d].sum()) for d in range(8)]
    for d in range(8):
        assert list(p.recv_off[d]) == [int(p.rows[:s, d].sum()) for s in range(8)]
    assert p.max_rows >= p.rows.sum(axis=0).max() and p.rows.sum(axis=0).max() > p.rows.sum(axis=0).mean()      # skewed
    u = uniform_plan(4, 10, 64)
    assert (u.rows == 10).all() and list(u.recv_off[2]) == [0, 10, 20, 30] and u.max_rows == 40


# ------------------------------------------------------------------------------------------------- the daemons as processes
def _root():
    import os
    return os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_scheduler_daemon_process_places_a_gang(api):
    """`python -m ...scheduler.daemon` as a process (flags, client construction, loop): a gated two-pod job ends up pinned to the
    two hosts that share a sub-block, and the gates are gone."""
    import os, subprocess, sys, time
    for i, lab in enumerate([("b1", "s1", "h1"), ("b2", "s9", "h7"), ("b1", "s1", "h2")]):
        n = node(f"node{i}", *lab)
        api.nodes[n["metadata"]["name"]] = n
    gate = "gke.io/topology-aware-auto-llm"
    for i in range(2):
        api.add_pod(pod(f"llm-{i}", job="llm", index=i, gate=gate))
    proc = subprocess.Popen([sys.executable, "-m", "container_engine_accelerators_b200.scheduler.daemon", "--kube-url", api.url, "--startup-cooloff", "0", "--gang-settle", "0",
                             "--gate-cooloff", "0", "--interval", "0.1", "--iterations", "3"], cwd=_root(), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    out, _ = proc.communicate(timeout=60)
    assert proc.returncode == 0, out
    placed = {}
    for i in range(2):
        spec = api.pods[("default", f"llm-{i}")]["spec"]
        assert spec["schedulingGates"] == [], out
        placed[i] = spec["affinity"]["nodeAffinity"]["requiredDuringSchedulingIgnoredDuringExecution"]["nodeSelectorTerms"][0]["matchExpressions"][0]["values"][0]
    assert sorted(placed.values()) == ["node0", "node2"]                      # the two hosts of block b1 / sub-block s1, not the far one
    assert testing.rbac_violations(_scheduler_role(), api.requests) == []


def test_labeler_process_with_a_fake_metadata_server(api):
    import subprocess, sys, threading
    from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

    class Meta(BaseHTTPRequestHandler):
        def log_message(self, *a):
            pass

        def do_GET(self):
            ok = self.headers.get("Metadata-Flavor") == "Google"
            body = {"/name": "gke-node-9", "/attributes/physical_host": "/cluster-x/rack-7/host-3"}.get(self.path)
            self.send_response(200 if ok and body else 404); self.end_headers()
            if ok and body:
                self.wfile.write(body.encode())
    srv = ThreadingHTTPServer(("127.0.0.1", 0), Meta)
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    try:
        api.add_node("gke-node-9", labels={"keep": "1"})
        r = subprocess.run([sys.executable, "-m", "container_engine_accelerators_b200.scheduler.labeler", "--once", "--kube-url", api.url,
                            "--metadata-url", f"http://127.0.0.1:{srv.server_address[1]}"], cwd=_root(), capture_output=True, text=True, timeout=60)
        assert r.returncode == 0, r.stderr
        assert api.nodes["gke-node-9"]["metadata"]["labels"] == {"keep": "1", "topology.gke.io/cluster": "cluster-x", "topology.gke.io/rack": "rack-7", "topology.gke.io/host": "host-3"}
        r = subprocess.run([sys.executable, "-m", "container_engine_accelerators_b200.scheduler.labeler", "--once", "--kube-url", api.url,
                            "--metadata-url", f"http://127.0.0.1:{srv.server_address[1]}/missing"], cwd=_root(), capture_output=True, text=True, timeout=60)
        assert r.returncode == 0 and "Node name not found" in r.stderr          # logged, the daemon keeps going
    finally:
        srv.shutdown(); srv.server_close()


def test_python_kube_client_over_tls_with_token_and_ca(tmp_path, monkeypatch):
    """KubeClient.from_env with B200_KUBE_URL / TOKEN_FILE / CA_FILE (the overrides the native client has): verified TLS, bearer token sent;
    a CA that does not cover the server is refused."""
    import requests
    certs = tmp_path / "pki"; certs.mkdir()
    cert, key = testing.make_self_signed_cert(str(certs))
    other = tmp_path / "other"; other.mkdir()
    other_cert, _ = testing.make_self_signed_cert(str(other))
    (tmp_path / "token").write_text("tok-123\n")
    a = testing.FakeKubeApi().start(tls_cert=cert, tls_key=key)
    try:
        a.add_node("n1")
        monkeypatch.setenv("B200_KUBE_URL", a.url); monkeypatch.setenv("B200_KUBE_TOKEN_FILE", str(tmp_path / "token")); monkeypatch.setenv("B200_KUBE_CA_FILE", cert)
        kc = kube.KubeClient.from_env()
        assert kc.get_node("n1")["metadata"]["name"] == "n1" and a.bearer_tokens[-1] == "Bearer tok-123"
        monkeypatch.setenv("B200_KUBE_CA_FILE", other_cert)
        with pytest.raises(requests.exceptions.SSLError):
            kube.KubeClient.from_env().get_node("n1")
        monkeypatch.delenv("B200_KUBE_URL"); monkeypatch.delenv("KUBERNETES_SERVICE_HOST", raising=False)
        with pytest.raises(kube.KubeError, match="not running in a cluster"):
            kube.KubeClient.from_env()
    finally:
        a.stop()
