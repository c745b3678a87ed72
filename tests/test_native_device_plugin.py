"""The native C++ device plugin (build/agent/b200-device-plugin) driven by the same conformance doubles as the Python agent:
real grpcio client <-> hand-written HTTP/2 + HPACK server, and the binary's own HTTP/2 client <-> a grpcio stub kubelet.
Scenarios mirror tests/test_device_plugin.py (reference shape: pkg/gpu/nvidia/beta_plugin_test.go:36-614)."""
import json
import os
import signal
import subprocess
import time
import urllib.request

import grpc
import pytest

from container_engine_accelerators_b200.agent import sharing, testing
from container_engine_accelerators_b200.agent.plugin import DevicePluginClient


class Native:
    def __init__(self, tmp_path, native_build, gpus=2, mig_parts=0, config=None, with_kubelet=True, extra_args=(), env=None, pci=None):
        self.root = str(tmp_path)
        self.dev = testing.make_fake_dev(self.root, gpus)
        self.proc = testing.make_fake_mig(self.root, self.dev, gpus, mig_parts) if mig_parts else str(tmp_path / "proc")
        self.plugin_dir = str(tmp_path / "dp"); os.makedirs(self.plugin_dir, exist_ok=True)
        self.kubelet = testing.KubeletStub(self.plugin_dir).start() if with_kubelet else None
        cfg = tmp_path / "gpu_config.json"
        if config is not None:
            cfg.write_text(config if isinstance(config, str) else json.dumps(config))
        self.endpoint = "nvidiaGPU-native.sock"
        self.log = open(tmp_path / "plugin.log", "w")
        e = {**os.environ, "B200AGENT_NVML_LIB": os.path.join(native_build, "libfake_nvml.so"), "FAKE_NVML_DEV_DIR": self.dev, **(env or {})}
        self.proc_handle = subprocess.Popen([os.path.join(native_build, "b200-device-plugin"), "-plugin-directory", self.plugin_dir, "-gpu-config", str(cfg), "--dev-directory", self.dev,
                                             "--proc-directory", self.proc, "--pci-root", pci or str(tmp_path / "nopci"), "--plugin-endpoint", self.endpoint, "--gpu-check-interval", "0.6",
                                             "--socket-check-interval", "0.1", *extra_args], env=e, stderr=self.log, stdout=self.log)
        self.client = None
        self.checked = False

    def connect(self, timeout=10):
        path = os.path.join(self.plugin_dir, self.endpoint)
        deadline = time.time() + timeout
        while not os.path.exists(path):
            assert time.time() < deadline and self.proc_handle.poll() is None, open(self.log.name).read()
            time.sleep(0.05)
        self.client = DevicePluginClient(path)
        self.client.wait_ready()
        return self.client

    def logs(self):
        self.log.flush()
        return open(self.log.name).read()

    def close(self):
        if self.client:
            self.client.close()
        if self.proc_handle.poll() is None:
            self.proc_handle.send_signal(signal.SIGTERM)
            try:
                self.proc_handle.wait(5)
            except subprocess.TimeoutExpired:
                self.proc_handle.kill()
        if self.kubelet:
            self.kubelet.stop()
        self.log.close()
        if os.environ.get("B200_NATIVE_SAN") and not self.checked:       # instrumented build: any report fails the test that produced it
            self.checked = True
            text = open(self.log.name).read()
            assert "Sanitizer" not in text and "runtime error:" not in text, text[-4000:]


@pytest.fixture
def native(tmp_path, native_build):
    made = []

    def make(**kw):
        n = Native(tmp_path, native_build, **kw)
        made.append(n)
        return n
    yield make
    for n in made:
        n.close()


def first_list(client):
    stream = client.list_and_watch()
    return stream, {d.ID: d for d in next(stream).devices}


def test_register_list_allocate(native):
    n = native()
    reg = n.kubelet.wait_registration()                       # the binary's own HTTP/2 client against a grpcio server
    assert (reg.version, reg.endpoint, reg.resource_name) == ("v1beta1", n.endpoint, "nvidia.com/gpu") and not reg.HasField("options")
    c = n.connect()
    opts = c.options()
    assert not opts.pre_start_required and not opts.get_preferred_allocation_available
    stream, devs = first_list(c)
    assert set(devs) == {"nvidia0", "nvidia1"} and all(d.health == "Healthy" for d in devs.values())
    cr = c.allocate(["nvidia0"]).container_responses[0]
    assert len(cr.devices) == 5 and len(cr.mounts) == 2
    assert cr.devices[0].host_path.endswith("/dev/nvidia0") and all(d.permissions == "mrw" and d.host_path == d.container_path for d in cr.devices)
    assert [(m.host_path, m.container_path, m.read_only) for m in cr.mounts] == [("/home/kubernetes/bin/nvidia", "/usr/local/nvidia", True), ("/home/kubernetes/bin/nvidia/vulkan/icd.d", "/etc/vulkan/icd.d", True)]
    assert dict(cr.envs) == {}
    resp = c.allocate(["nvidia0", "nvidia1"], ["nvidia1"])
    assert [len(r.devices) for r in resp.container_responses] == [6, 5]
    with pytest.raises(grpc.RpcError) as ei:
        c.allocate(["nvidia9"])
    assert ei.value.code() == grpc.StatusCode.UNKNOWN and "invalid allocation request with non-existing device nvidia9" in ei.value.details()
    stream.cancel()


def test_many_sequential_calls_reuse_one_connection(native):
    """HPACK dynamic-table state must survive across requests on one HTTP/2 connection (grpcio indexes repeated headers)."""
    c = native().connect()
    for i in range(50):
        assert len(c.allocate([f"nvidia{i % 2}"]).container_responses[0].devices) == 5


def test_numa_topology_is_advertised(native, tmp_path):
    pci = testing.make_fake_pci(str(tmp_path), "0000:1b:00.0", 1)
    c = native(gpus=1, pci=pci).connect()
    _, devs = first_list(c)
    assert [n.ID for n in devs["nvidia0"].topology.nodes] == [1]


def test_time_sharing_and_mig(native, tmp_path):
    n = native(config={"GPUSharingConfig": {"GPUSharingStrategy": "time-sharing", "MaxSharedClientsPerGPU": 3}})
    c = n.connect()
    _, devs = first_list(c)
    assert set(devs) == {f"nvidia{g}/vgpu{k}" for g in range(2) for k in range(3)}
    assert c.allocate(["nvidia1/vgpu2"]).container_responses[0].devices[0].host_path.endswith("/dev/nvidia1")
    with pytest.raises(grpc.RpcError) as ei:
        c.allocate(["nvidia0/vgpu0", "nvidia0/vgpu1"])
    assert sharing.ERR_TIME_SHARING in ei.value.details()


def test_mig_seven_slices(native):
    c = native(gpus=1, mig_parts=7, config={"GPUPartitionSize": "1g.23gb"}).connect()
    _, devs = first_list(c)
    assert set(devs) == {f"nvidia0/gi{i}" for i in range(1, 8)}
    cr = c.allocate(["nvidia0/gi3"]).container_responses[0]
    assert len(cr.devices) == 7 and "/nvidia-caps/nvidia-cap" in cr.devices[1].host_path
    with pytest.raises(grpc.RpcError) as ei:
        c.allocate(["nvidia0/gi9"])
    assert "invalid allocation request with non-existing GPU partition: nvidia0/gi9" in ei.value.details()


def test_bad_config_falls_back_and_transport_profile(native):
    n = native(config="{broken json")
    c = n.connect()
    _, devs = first_list(c)
    assert set(devs) == {"nvidia0", "nvidia1"} and "falling back to default GPU config" in n.logs()
    n.close()           # same plugin dir and endpoint name: the first instance must be gone (it unlinks the socket on exit)
    n2 = native(config={"Transport": {"Name": "b200coll", "Env": {"B200COLL_ALGO": "nvls"}}}, with_kubelet=False)
    env = dict(n2.connect().allocate(["nvidia0"]).container_responses[0].envs)
    assert env["B200COLL_LIB"] == "/usr/local/nvidia/lib64/libb200coll.so" and env["B200COLL_ALGO"] == "nvls" and env["LD_LIBRARY_PATH"] == "/usr/local/nvidia/lib64"


def test_hot_add_and_socket_removal_restart(native):
    n = native()
    n.kubelet.wait_registration()
    c = n.connect()
    stream, devs = first_list(c)
    stream.cancel()
    testing.add_fake_gpu(n.dev, 2)
    assert n.kubelet.wait_registration(timeout=10).endpoint == n.endpoint
    c.close()
    c = n.connect()
    _, devs = first_list(c)
    assert set(devs) == {"nvidia0", "nvidia1", "nvidia2"}
    c.close(); n.client = None
    os.unlink(os.path.join(n.plugin_dir, n.endpoint))
    assert n.kubelet.wait_registration(timeout=10).endpoint == n.endpoint


def test_xid_event_streams_unhealthy_and_blocks_allocation(native, tmp_path):
    events = tmp_path / "events.txt"
    events.write_text("")
    n = native(extra_args=["-enable-health-monitoring"], env={"FAKE_NVML_EVENTS": str(events), "XID_CONFIG": "31"})
    c = n.connect()
    stream = c.list_and_watch()
    assert {d.ID: d.health for d in next(stream).devices} == {"nvidia0": "Healthy", "nvidia1": "Healthy"}
    time.sleep(0.5)
    events.write_text("0 13\n1 31\n")                 # 13 is not critical; 31 is (XID_CONFIG)
    assert {d.ID: d.health for d in next(stream).devices} == {"nvidia0": "Healthy", "nvidia1": "Unhealthy"}
    with pytest.raises(grpc.RpcError) as ei:
        c.allocate(["nvidia1"])
    assert "invalid allocation request with unhealthy device nvidia1" in ei.value.details()
    stream.cancel()


@pytest.mark.parametrize("service", ["v1alpha1.PodResourcesLister", "v1.PodResourcesLister"])
def test_metrics_endpoint(native, tmp_path, service):
    from tests.test_health_metrics import PodResourcesStub
    sock = str(tmp_path / "pr.sock")
    stub = PodResourcesStub(sock, [("default", "p1", "c1", "nvidia.com/gpu", ["nvidia0"]), ("default", "p2", "c1", "nvidia.com/gpu", ["nvidia1/vgpu0"])], service=service)
    import socket as _s
    s = _s.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    try:
        import struct
        shm = tmp_path / "shm"; shm.mkdir()
        page = bytearray(4096); page[0:8] = b"B200COLL"
        struct.pack_into("<6I", page, 8, 2, 4242, 3, 8, 3, 1)
        struct.pack_into("<21Q", page, 64, 10, 0, 0, 2, 7, 1, 1 << 30, 0, 0, 4096, 512, 64, 0, 5, 0, 2, 13, 0, 0, 21, 0)
        struct.pack_into("<3Q", page, 64 + 21 * 8, 4, 5, 6000)                  # point-to-point: sends, recvs, bytes
        struct.pack_into("<6Q", page, 64 + 24 * 8, 9, 1 << 32, 3, 2, 40, 41)    # host path: calls / bytes / zero-copy / pipelined; bulk and generic launches
        struct.pack_into("<Q", page, 32, int(time.time()))                      # freshly updated
        (shm / "b200coll.4242.3").write_bytes(page)
        stale = bytearray(page); struct.pack_into("<6I", stale, 8, 2, 999, 0, 8, 0, 1); struct.pack_into("<Q", stale, 32, int(time.time()) - 7200)
        (shm / "b200coll.999.0").write_bytes(stale)                            # a dead process's leftover: not exported
        n = native(extra_args=["-enable-container-gpu-metrics", "-gpu-metrics-port", str(port), "-gpu-metrics-collection-interval", "200", "--pod-resources-socket", sock,
                               "--coll-stats-dir", str(shm)], env={"FAKE_NVML_UTIL": "40,60,80"})
        n.connect()
        deadline = time.time() + 10
        body = ""
        while time.time() < deadline:
            try:
                body = urllib.request.urlopen(f"http://127.0.0.1:{port}/metrics", timeout=2).read().decode()
                if "duty_cycle{" in body:
                    break
            except OSError:
                pass
            time.sleep(0.2)
        assert 'duty_cycle{namespace="default",pod="p1",container="c1",make="nvidia",accelerator_id="GPU-fake-0",model="NVIDIA B200"} 60' in body
        assert 'request{namespace="default",pod="p1",container="c1",resource_name="nvidia.com/gpu"} 1' in body
        assert 'request{namespace="default",pod="p2",container="c1",resource_name="nvidia.com/gpu"} 0' in body            # virtual ids dropped
        assert 'memory_total_gpu_node{make="nvidia",accelerator_id="GPU-fake-1",model="NVIDIA B200"} ' in body
        assert 'b200coll_calls{pid="4242",rank="3",op="all_reduce"} 10' in body and 'b200coll_calls{pid="4242",rank="3",op="broadcast"} 7' in body
        assert f'b200coll_bytes{{pid="4242",rank="3",op="all_reduce"}} {1 << 30}' in body
        assert 'b200coll_algo_calls{pid="4242",rank="3",algo="nvls"} 13' in body
        assert 'b200coll_p2p_calls{pid="4242",rank="3",dir="send"} 4' in body and 'b200coll_p2p_bytes{pid="4242",rank="3"} 6000' in body
        assert 'b200coll_host_calls{pid="4242",rank="3",path="total"} 9' in body and 'b200coll_host_calls{pid="4242",rank="3",path="pipelined"} 2' in body
        assert f'b200coll_host_bytes{{pid="4242",rank="3"}} {1 << 32}' in body
        assert 'b200coll_kernel_family_launches{pid="4242",rank="3",family="bulk"} 40' in body and 'family="generic"} 41' in body
        assert 'pid="999"' not in body
    finally:
        stub.server.stop(0)


# ------------------------------------------------------------------------------------------------- Kubernetes side effects
def _wait(pred, timeout=10.0, what=""):
    deadline = time.time() + timeout
    while time.time() < deadline:
        if pred():
            return
        time.sleep(0.05)
    raise AssertionError(f"timed out waiting for {what}")


def _device_plugin_role():
    import yaml
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return next(d for d in yaml.safe_load_all(open(os.path.join(root, "deploy", "device-plugin", "rbac.yaml"))) if d["kind"] == "ClusterRole")


def _xid_condition(api, node="node-a"):
    return next((c for c in api.nodes[node]["status"]["conditions"] if c["type"] == "XidCriticalError"), None)


def test_xid_event_condition_and_heartbeat_through_kube_api(native, tmp_path):
    """health_check/health_checker.go:288-358,395-449: Event per critical Xid, Node condition whose reason is the JSON set
    of Xids and whose message is the boot id, heartbeat refresh; non-monitored Xids only produce the Event."""
    api = testing.FakeKubeApi().start()
    try:
        api.add_node("node-a", boot_id="boot-7")
        events = tmp_path / "events.txt"; events.write_text("")
        n = native(extra_args=["-enable-health-monitoring", "--xid-heartbeat-interval", "0.2"],
                   env={"FAKE_NVML_EVENTS": str(events), "B200_KUBE_URL": api.url, "NODE_NAME": "node-a", "XID_CONFIG": "31"})
        c = n.connect()
        stream = c.list_and_watch(); next(stream)
        time.sleep(0.5)
        events.write_text("0 79\n")                       # monitored (condition) but not health-critical here
        _wait(lambda: _xid_condition(api) is not None, what="XidCriticalError condition")
        cond = _xid_condition(api)
        assert cond["status"] == "True" and json.loads(cond["reason"]) == {"79": True} and cond["message"] == "boot-7"
        assert [e["message"] for e in api.events] == ["Caught XID error, XID=79"]
        ev = api.events[0]
        assert ev["involvedObject"] == {"kind": "Node", "name": "node-a", "uid": "uid-node-a", "apiVersion": "v1"}
        assert ev["type"] == "Warning" and ev["reason"] == "XIDError" and ev["source"] == {"component": "nvidia-gpu-device-plugin"}
        assert api.nodes["node-a"]["status"]["nodeInfo"]["bootID"] == "boot-7"          # the rest of the status survived the round trip
        assert any(c2["type"] == "Ready" for c2 in api.nodes["node-a"]["status"]["conditions"])
        first_beat = cond["lastHeartbeatTime"]
        with open(events, "a") as f:                      # the fake NVML reads the file as an append-only log
            f.write("0 79\n1 48\n1 31\n")                # 79 again (dedup), 48 (always critical + monitored), 31 (critical via XID_CONFIG, not monitored)
        assert {d.ID: d.health for d in next(stream).devices} == {"nvidia0": "Healthy", "nvidia1": "Unhealthy"}
        _wait(lambda: len(api.events) == 4, what="4 events")
        _wait(lambda: json.loads(_xid_condition(api)["reason"]) == {"48": True, "79": True}, what="reason set {48,79}")
        assert "already includes this XID 79" in n.logs()
        time.sleep(1.3)                                   # RFC 3339 stamps have 1 s resolution
        _wait(lambda: _xid_condition(api)["lastHeartbeatTime"] > first_beat, timeout=5, what="heartbeat refresh")
        assert ("PUT", "/api/v1/nodes/node-a/status", "application/json") in api.requests
        assert testing.rbac_violations(_device_plugin_role(), api.requests) == []       # everything it did is covered by deploy/device-plugin/rbac.yaml
        stream.cancel()
    finally:
        api.stop()


def test_stale_xid_condition_is_cleared_after_reboot_and_kept_otherwise(native, tmp_path):
    """health_checker.go:129-160: boot id differs from the one recorded in the condition -> repaired -> remove it."""
    api = testing.FakeKubeApi().start()
    try:
        stale = {"type": "XidCriticalError", "status": "True", "reason": "{\"48\":true}", "message": "boot-old"}
        api.add_node("node-a", boot_id="boot-new", conditions=[{"type": "Ready", "status": "True"}, dict(stale)])
        api.add_node("node-b", boot_id="boot-old", conditions=[{"type": "Ready", "status": "True"}, dict(stale)])
        api.fail_next_gets = 1                            # first GET fails: the reset retries with backoff
        n = native(extra_args=["-enable-health-monitoring"], env={"B200_KUBE_URL": api.url, "NODE_NAME": "node-a"})
        n.connect()
        _wait(lambda: _xid_condition(api, "node-a") is None, timeout=8, what="stale condition removal")
        _wait(lambda: "Successfully removed XIDCriticalError condition from node node-a." in n.logs(), what="removal log line")
        assert [c["type"] for c in api.nodes["node-a"]["status"]["conditions"]] == ["Ready"]
        n.close()
        n2 = Native(tmp_path / "b", os.path.dirname(n.proc_handle.args[0]), extra_args=["-enable-health-monitoring"], env={"B200_KUBE_URL": api.url, "NODE_NAME": "node-b"})
        try:
            n2.connect()
            _wait(lambda: "XIDCriticalError condition doesn't exist for node node-b." in n2.logs(), what="no-op reset")
            assert _xid_condition(api, "node-b") is not None      # same boot: the fault is still current
        finally:
            n2.close()
    finally:
        api.stop()


def test_driver_version_annotations_by_server_side_apply(native):
    """version_visibility.go:38-86: four annotations, apply patch with fieldManager gpu-device-plugin and force."""
    api = testing.FakeKubeApi().start()
    try:
        api.add_node("node-a", annotations={"keep": "me"})
        n = native(extra_args=["--publish-driver-version"], env={"B200_KUBE_URL": api.url, "NODE_NAME": "node-a", "FAKE_NVML_DRIVER": "570.124.06"})
        n.connect()
        _wait(lambda: "cloud.google.com/cuda.driver-version.full" in api.nodes["node-a"]["metadata"]["annotations"], what="annotations")
        ann = api.nodes["node-a"]["metadata"]["annotations"]
        assert ann == {"keep": "me", "cloud.google.com/cuda.driver-version.major": "570", "cloud.google.com/cuda.driver-version.minor": "124",
                       "cloud.google.com/cuda.driver-version.revision": "06", "cloud.google.com/cuda.driver-version.full": "570.124.06"}
        assert ("PATCH", "/api/v1/nodes/node-a", "application/apply-patch+yaml") in api.requests
        assert testing.rbac_violations(_device_plugin_role(), api.requests) == []
    finally:
        api.stop()


def test_malformed_driver_version_is_not_published(native):
    api = testing.FakeKubeApi().start()
    try:
        api.add_node("node-a")
        n = native(extra_args=["--publish-driver-version"], env={"B200_KUBE_URL": api.url, "NODE_NAME": "node-a", "FAKE_NVML_DRIVER": "570.x"})
        n.connect()
        _wait(lambda: "unexpected driver version format: 570.x" in n.logs(), what="format error")
        assert api.nodes["node-a"]["metadata"]["annotations"] == {}
    finally:
        api.stop()


def test_kube_client_over_tls_verifies_the_server(native, tmp_path):
    """The in-cluster path: https + bearer token + CA bundle. A certificate the CA bundle does not cover must be refused."""
    certs = tmp_path / "pki"; certs.mkdir()
    cert, key = testing.make_self_signed_cert(str(certs))
    other = tmp_path / "other"; other.mkdir()
    other_cert, _ = testing.make_self_signed_cert(str(other))
    token = tmp_path / "token"; token.write_text("s3cr3t\n")
    api = testing.FakeKubeApi().start(tls_cert=cert, tls_key=key)
    try:
        assert api.url.startswith("https://")
        api.add_node("node-a")
        env = {"B200_KUBE_URL": api.url, "B200_KUBE_TOKEN_FILE": str(token), "NODE_NAME": "node-a", "FAKE_NVML_DRIVER": "580.159.03"}
        n = native(extra_args=["--publish-driver-version"], env={**env, "B200_KUBE_CA_FILE": cert})
        n.connect()
        _wait(lambda: api.nodes["node-a"]["metadata"]["annotations"].get("cloud.google.com/cuda.driver-version.major") == "580", what="annotations over TLS")
        assert api.bearer_tokens[-1] == "Bearer s3cr3t"
        n.close()
        # the same server reached by name: SNI + host-name verification against the certificate's DNS subject-alt-name
        api.nodes["node-a"]["metadata"]["annotations"].clear()
        by_name = Native(tmp_path / "c", os.path.dirname(n.proc_handle.args[0]), extra_args=["--publish-driver-version"],
                         env={**env, "B200_KUBE_URL": api.url.replace("127.0.0.1", "localhost"), "B200_KUBE_CA_FILE": cert})
        try:
            by_name.connect()
            _wait(lambda: api.nodes["node-a"]["metadata"]["annotations"].get("cloud.google.com/cuda.driver-version.major") == "580", what="annotations over TLS by host name")
        finally:
            by_name.close()
        api.nodes["node-a"]["metadata"]["annotations"].clear()
        n2 = Native(tmp_path / "b", os.path.dirname(n.proc_handle.args[0]), extra_args=["--publish-driver-version"], env={**env, "B200_KUBE_CA_FILE": other_cert})
        try:
            n2.connect()
            _wait(lambda: "TLS handshake with 127.0.0.1 failed" in n2.logs(), what="handshake refusal")
            assert api.nodes["node-a"]["metadata"]["annotations"] == {}
        finally:
            n2.close()
    finally:
        api.stop()


def test_health_monitoring_without_a_cluster_still_reports_devices(native, tmp_path):
    events = tmp_path / "events.txt"; events.write_text("")
    n = native(extra_args=["-enable-health-monitoring"], env={"FAKE_NVML_EVENTS": str(events), "KUBERNETES_SERVICE_HOST": "", "B200_KUBE_URL": ""})
    c = n.connect()
    stream = c.list_and_watch(); next(stream)
    time.sleep(0.5)
    events.write_text("0 48\n")
    assert {d.ID: d.health for d in next(stream).devices} == {"nvidia0": "Unhealthy", "nvidia1": "Healthy"}
    assert "failed to build kube client: not running in a cluster" in n.logs()
    stream.cancel()


# ------------------------------------------------------------------------------------------------- hostile HTTP/2 peers
def _frame(ftype, flags, stream, payload=b""):
    return len(payload).to_bytes(3, "big") + bytes([ftype, flags]) + stream.to_bytes(4, "big") + payload


def _hpack_literal(headers):
    out = b""
    for k, v in headers:
        out += b"\x00" + bytes([len(k)]) + k.encode() + bytes([len(v)]) + v.encode()
    return out


def test_malformed_http2_traffic_does_not_take_the_plugin_down(native):
    """A privileged daemon must shrug off garbage on its socket: every abusive connection is dropped or answered with an error and
    the next well-formed kubelet call still works (run under ASan/TSan with B200_NATIVE_SAN to catch memory errors, not just crashes)."""
    import random
    import socket
    n = native()
    c = n.connect()
    path = os.path.join(n.plugin_dir, n.endpoint)
    preface = b"PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n"
    req_headers = _hpack_literal([(":method", "POST"), (":scheme", "http"), (":path", "/v1beta1.DevicePlugin/ListAndWatch"), ("content-type", "application/grpc"), ("te", "trailers")])
    grpc_empty = b"\x00\x00\x00\x00\x00"
    rng = random.Random(7)
    attacks = [
        b"GET / HTTP/1.1\r\nHost: x\r\n\r\n",                                                   # not HTTP/2 at all
        preface[:10],                                                                           # truncated preface, then close
        preface + b"\xff\xff\xff\x00\x00\x00\x00\x00\x01",                                     # 16 MiB frame announced
        preface + _frame(1, 0x4, 1, b"\x80"),                                                   # HPACK index 0
        preface + _frame(1, 0x4, 1, b"\x00\x7f\xff\xff\xff\xff\xff\xff\xff\xff\x7f"),           # HPACK string length that would wrap
        preface + _frame(1, 0x4, 0, req_headers),                                               # HEADERS on stream 0
        preface + _frame(1, 0x4, 2, req_headers),                                               # even (server-initiated) stream id
        preface + _frame(1, 0x4 | 0x8, 1, b"\xff" + req_headers),                               # pad length > payload
        preface + _frame(9, 0x4, 5, req_headers),                                               # CONTINUATION without HEADERS
        preface + _frame(1, 0x0, 1, req_headers[:7]) + b"".join(_frame(9, 0, 1, b"A" * 16000) for _ in range(6)),      # endless header block
        preface + _frame(1, 0x4, 1, req_headers) + _frame(0, 0x1, 1, grpc_empty) + _frame(0, 0x1, 1, grpc_empty)       # END_STREAM twice on a streaming RPC
        + _frame(1, 0x4, 1, req_headers),                                                                                # ... and the stream id reused
        preface + _frame(1, 0x4, 1, req_headers) + _frame(0, 0x0, 1, b"\x00\xff\xff\xff\xff" + b"B" * 100) + _frame(0, 0x1, 1, b""),   # gRPC length prefix of 4 GiB
        preface + _frame(8, 0, 0, b"\x7f\xff\xff\xff") * 4 + _frame(4, 0, 0, b"\x00\x04\xff\xff\xff\xff") + _frame(3, 0, 9, b"\x00\x00\x00\x08"),  # window overflow, huge initial window, RST of unknown stream
        preface + bytes(rng.getrandbits(8) for _ in range(4096)),                               # noise after a valid preface
    ]
    for raw in attacks:
        s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        s.settimeout(0.5)
        s.connect(path)
        try:
            s.sendall(raw)
            try:
                while s.recv(65536):
                    pass                                      # drain whatever it answers until it closes or goes quiet
            except (socket.timeout, ConnectionResetError):
                pass
        except BrokenPipeError:
            pass
        finally:
            s.close()
        assert n.proc_handle.poll() is None, n.logs()[-2000:]
    many = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)   # more concurrent streams than the server admits: the surplus is refused, not served
    many.settimeout(1); many.connect(path)
    many.sendall(preface + b"".join(_frame(1, 0x4, 2 * i + 1, req_headers) for i in range(300)))
    got = b""
    try:
        while True:
            chunk = many.recv(65536)
            if not chunk:
                break
            got += chunk
    except socket.timeout:
        pass
    many.close()
    assert _frame(3, 0, 2 * 299 + 1, (7).to_bytes(4, "big")) in got           # RST_STREAM(REFUSED_STREAM) for the last one
    # and the kubelet's own connection, opened before the abuse, still works, as does a fresh one
    assert len(c.allocate(["nvidia0"]).container_responses[0].devices) == 5
    c2 = DevicePluginClient(path); c2.wait_ready()
    stream = c2.list_and_watch()
    assert {d.ID for d in next(stream).devices} == {"nvidia0", "nvidia1"}
    stream.cancel(); c2.close()


# ------------------------------------------------------------------------------------------------- preferred allocation (opt-in)
def test_preferred_allocation_native_matches_the_python_policy(native, tmp_path):
    """-preferred-allocation-policy: options advertised in Register and GetDevicePluginOptions, NUMA-aligned answers; the C++ and
    Python implementations are compared on the same requests (agent/preferred.py is the specification)."""
    from container_engine_accelerators_b200.agent import preferred
    pci = None
    for i, node in enumerate([0, 0, 1, 1]):                       # fake NVML bus ids are 00000000:1B+i:00.0
        pci = testing.make_fake_pci(str(tmp_path), f"0000:{0x1b + i:02x}:00.0", node)
    n = native(gpus=4, pci=pci, extra_args=["-preferred-allocation-policy", "spread"])
    reg = n.kubelet.wait_registration()
    assert reg.options.get_preferred_allocation_available and not reg.options.pre_start_required
    c = n.connect()
    assert c.options().get_preferred_allocation_available
    numa = {"nvidia0": 0, "nvidia1": 0, "nvidia2": 1, "nvidia3": 1}.get
    cases = [(["nvidia0", "nvidia2", "nvidia3"], [], 2), (["nvidia0", "nvidia1", "nvidia2", "nvidia3"], ["nvidia3"], 2), (["nvidia3", "nvidia1", "nvidia0"], [], 3),
             (["nvidia1"], [], 2), (["nvidia0", "nvidia1", "nvidia2", "nvidia3"], [], 4), (["nvidia2", "nvidia0"], ["nvidia0", "nvidia0"], 1)]
    for available, must, size in cases:
        assert c.preferred(available, must, size) == preferred.preferred_allocation(available, must, size, numa, "spread"), (available, must, size)
    assert c.preferred(["nvidia0", "nvidia2", "nvidia3"], [], 2) == ["nvidia2", "nvidia3"]
    import random
    rng = random.Random(11)
    ids = ["nvidia0", "nvidia1", "nvidia2", "nvidia3"]
    for _ in range(60):                                            # differential check on random requests
        available = rng.sample(ids, rng.randint(0, 4))
        must = rng.sample(available, rng.randint(0, min(2, len(available))))
        size = rng.randint(len(must), 4)
        assert c.preferred(available, must, size) == preferred.preferred_allocation(available, must, size, numa, "spread"), (available, must, size)


@pytest.mark.parametrize("policy", ["spread", "packed"])
def test_preferred_allocation_native_shared_gpus(native, policy):
    from container_engine_accelerators_b200.agent import preferred
    n = native(config={"GPUSharingConfig": {"GPUSharingStrategy": "time-sharing", "MaxSharedClientsPerGPU": 3}}, extra_args=["-preferred-allocation-policy", policy])
    c = n.connect()
    free = ["nvidia0/vgpu2", "nvidia1/vgpu0", "nvidia1/vgpu1", "nvidia1/vgpu2"]        # nvidia0 is busy (1 replica left), nvidia1 idle
    for size in (1, 2, 3):
        assert c.preferred(free, [], size) == preferred.preferred_allocation(free, [], size, lambda d: None, policy), (policy, size)
    assert c.preferred(free, [], 1) == (["nvidia1/vgpu0"] if policy == "spread" else ["nvidia0/vgpu2"])


def test_native_without_the_policy_flag_keeps_the_reference_contract(native):
    n = native()
    assert not n.kubelet.wait_registration().HasField("options")
    c = n.connect()
    assert not c.options().get_preferred_allocation_available and c.preferred(["nvidia0"], [], 1) == []
    assert "GetPreferredAllocation should NOT be called" in n.logs()


def test_native_status_updates_survive_write_conflicts(native, tmp_path):
    """409 on the status PUT (somebody wrote the Node between our GET and PUT): the native plugin re-reads and retries."""
    api = testing.FakeKubeApi().start()
    try:
        api.add_node("node-a", boot_id="boot-7")
        events = tmp_path / "events.txt"; events.write_text("")
        n = native(extra_args=["-enable-health-monitoring", "--xid-heartbeat-interval", "30"], env={"FAKE_NVML_EVENTS": str(events), "B200_KUBE_URL": api.url, "NODE_NAME": "node-a"})
        c = n.connect()
        stream = c.list_and_watch(); next(stream)
        time.sleep(0.5)
        api.stale_next_puts = 2
        with open(events, "a") as f:
            f.write("0 79\n")
        _wait(lambda: _xid_condition(api) is not None, what="condition despite two conflicts")
        assert json.loads(_xid_condition(api)["reason"]) == {"79": True} and api.conflicts == 2
        assert "retrying the status update" in n.logs()
        stream.cancel()
    finally:
        api.stop()
