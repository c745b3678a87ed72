"""DevicePlugin Register / ListAndWatch / Allocate against a stub kubelet over real Unix sockets (BASELINE config #1).
Shape follows the reference's in-process integration test (pkg/gpu/nvidia/beta_plugin_test.go:36-614): four modes
(plain, time-sharing, MIG, MIG + time-sharing), 5 device specs + 2 mounts for one GPU, 7 specs for one MIG slice,
unknown id must fail the RPC, hot-added GPU is re-advertised after the rescan interval."""
import threading
import time

import grpc
import pytest

from container_engine_accelerators_b200.agent import manager as mgr
from container_engine_accelerators_b200.agent import nvml, preferred, protos, sharing, testing
from container_engine_accelerators_b200.agent.config import GPUConfig, GPUSharingConfig, TransportConfig
from container_engine_accelerators_b200.agent.plugin import DevicePluginClient

MOUNTS = [mgr.Mount("/home/kubernetes/bin/nvidia", "/usr/local/nvidia", True), mgr.Mount("/home/kubernetes/bin/nvidia/vulkan/icd.d", "/etc/vulkan/icd.d", True)]


class Harness:
    def __init__(self, tmp_path, cfg: GPUConfig, gpus=2, mig_parts=0, with_kubelet=True, policy="none", numa=None):
        self.root = str(tmp_path)
        self.dev = testing.make_fake_dev(self.root, gpus)
        self.proc = testing.make_fake_mig(self.root, self.dev, gpus, mig_parts) if mig_parts else str(tmp_path / "proc")
        self.plugin_dir = str(tmp_path / "device-plugin")
        import os
        os.makedirs(self.plugin_dir, exist_ok=True)
        self.kubelet = testing.KubeletStub(self.plugin_dir).start() if with_kubelet else None
        cfg.add_defaults_and_validate()
        pci_root, bus_ids = nvml.PCI_DEVICES_ROOT, None
        if numa:                                    # [numa node per gpu index]: fake sysfs entries, one bus id per GPU
            bus_ids = [f"0000:{0x1b + 0x10 * i:02x}:00.0" for i in range(gpus)]
            for i, node in enumerate(numa):
                pci_root = testing.make_fake_pci(self.root, bus_ids[i], node)
        self.ngm = mgr.GPUManager(self.dev, self.proc, list(MOUNTS), cfg, nvml=nvml.MockNvml(self.dev, bus_ids=bus_ids), pci_root=pci_root, gpu_check_interval=0.6,
                                  socket_check_interval=0.1, preferred_allocation_policy=policy)
        self.ngm.start()
        self.endpoint = "nvidiaGPU-test.sock"
        self.thread = threading.Thread(target=self.ngm.serve, args=(self.plugin_dir, "kubelet.sock", self.endpoint), daemon=True)
        self.thread.start()
        self.client = None

    def connect(self):
        import os
        assert self.ngm.serving.wait(10), "plugin never started serving"
        self.client = DevicePluginClient(os.path.join(self.plugin_dir, self.endpoint))
        self.client.wait_ready()
        return self.client

    def close(self):
        if self.client:
            self.client.close()
        self.ngm.stop()
        self.thread.join(5)
        if self.kubelet:
            self.kubelet.stop()


@pytest.fixture
def harness(tmp_path):
    made = []

    def make(cfg=None, **kw):
        h = Harness(tmp_path, cfg or GPUConfig(), **kw)
        made.append(h)
        return h
    yield make
    for h in made:
        h.close()


def first_list(client):
    stream = client.list_and_watch()
    resp = next(stream)
    return stream, {d.ID: d for d in resp.devices}


def test_register_list_allocate_plain(harness):
    h = harness()
    reg = h.kubelet.wait_registration()
    assert (reg.version, reg.endpoint, reg.resource_name) == ("v1beta1", h.endpoint, "nvidia.com/gpu")
    assert not reg.HasField("options")               # no Options: the kubelet must never call PreStart/GetPreferredAllocation
    c = h.connect()
    opts = c.options()
    assert not opts.pre_start_required and not opts.get_preferred_allocation_available
    stream, devs = first_list(c)
    assert set(devs) == {"nvidia0", "nvidia1"}
    assert all(d.health == "Healthy" for d in devs.values())
    resp = c.allocate(["nvidia0"])
    cr = resp.container_responses[0]
    assert len(cr.devices) == 5 and len(cr.mounts) == 2          # reference beta_plugin_test.go:354-359
    paths = [d.host_path for d in cr.devices]
    assert paths[0].endswith("/dev/nvidia0")
    assert {p.rsplit("/", 1)[1] for p in paths[1:]} == {"nvidiactl", "nvidia-uvm", "nvidia-uvm-tools", "nvidia-modeset"}
    assert all(d.permissions == "mrw" and d.host_path == d.container_path for d in cr.devices)
    assert [(m.host_path, m.container_path, m.read_only) for m in cr.mounts] == [(m.host_path, m.container_path, True) for m in MOUNTS]
    assert dict(cr.envs) == {}
    with pytest.raises(grpc.RpcError) as ei:
        c.allocate(["nvidia9"])
    assert "invalid allocation request with non-existing device nvidia9" in ei.value.details()
    stream.cancel()


def test_multi_container_and_multi_device(harness):
    h = harness()
    c = h.connect()
    resp = c.allocate(["nvidia0", "nvidia1"], ["nvidia1"])
    assert [len(r.devices) for r in resp.container_responses] == [6, 5]


def test_hot_add_gpu_is_rediscovered(harness):
    h = harness()
    h.kubelet.wait_registration()
    c = h.connect()
    stream, devs = first_list(c)
    assert set(devs) == {"nvidia0", "nvidia1"}
    stream.cancel()
    testing.add_fake_gpu(h.dev, 2)
    reg = h.kubelet.wait_registration(timeout=10)     # server restarts and re-registers after the rescan interval
    assert reg.endpoint == h.endpoint
    c.close()
    c = h.connect()
    _, devs = first_list(c)
    assert set(devs) == {"nvidia0", "nvidia1", "nvidia2"}


def test_health_update_is_streamed_and_blocks_allocation(harness):
    h = harness()
    c = h.connect()
    stream = c.list_and_watch()
    assert {d.ID: d.health for d in next(stream).devices} == {"nvidia0": "Healthy", "nvidia1": "Healthy"}
    from container_engine_accelerators_b200.agent.mig import Device
    assert h.ngm.report_unhealthy(Device("nvidia1", protos.UNHEALTHY))
    upd = {d.ID: d.health for d in next(stream).devices}
    assert upd == {"nvidia0": "Healthy", "nvidia1": "Unhealthy"}
    with pytest.raises(grpc.RpcError) as ei:
        c.allocate(["nvidia1"])
    assert "invalid allocation request with unhealthy device nvidia1" in ei.value.details()
    stream.cancel()


def test_time_sharing(harness):
    h = harness(GPUConfig(sharing=GPUSharingConfig(sharing.TIME_SHARING, 3)))
    c = h.connect()
    _, devs = first_list(c)
    assert set(devs) == {f"nvidia{g}/vgpu{k}" for g in range(2) for k in range(3)}
    cr = c.allocate(["nvidia1/vgpu2"]).container_responses[0]
    assert len(cr.devices) == 5 and cr.devices[0].host_path.endswith("/dev/nvidia1")
    with pytest.raises(grpc.RpcError) as ei:
        c.allocate(["nvidia0/vgpu0", "nvidia0/vgpu1"])
    assert sharing.ERR_TIME_SHARING in ei.value.details()


def test_mig_seven_slices(harness):
    """BASELINE config #5: 1g.23gb x7 on B200 -> 7 resources per GPU, 3 specs + defaults per slice."""
    h = harness(GPUConfig(gpu_partition_size="1g.23gb"), gpus=1, mig_parts=7)
    c = h.connect()
    _, devs = first_list(c)
    assert set(devs) == {f"nvidia0/gi{i}" for i in range(1, 8)}
    cr = c.allocate(["nvidia0/gi3"]).container_responses[0]
    assert len(cr.devices) == 7 and len(cr.mounts) == 2           # reference beta_plugin_test.go:535
    assert cr.devices[0].host_path.endswith("/dev/nvidia0")
    assert "/nvidia-caps/nvidia-cap" in cr.devices[1].host_path and "/nvidia-caps/nvidia-cap" in cr.devices[2].host_path
    with pytest.raises(grpc.RpcError) as ei:
        c.allocate(["nvidia0/gi9"])
    assert "invalid allocation request with non-existing GPU partition: nvidia0/gi9" in ei.value.details()


def test_mig_with_time_sharing(harness):
    h = harness(GPUConfig(gpu_partition_size="3g.90gb", sharing=GPUSharingConfig(sharing.TIME_SHARING, 2)), gpus=2, mig_parts=2)
    c = h.connect()
    _, devs = first_list(c)
    assert len(devs) == 2 * 2 * 2 and "nvidia1/gi2/vgpu1" in devs
    cr = c.allocate(["nvidia1/gi2/vgpu1"]).container_responses[0]
    assert len(cr.devices) == 7


def test_mig_wrong_partition_count_fails_start(tmp_path):
    cfg = GPUConfig(gpu_partition_size="1g.23gb")
    cfg.add_defaults_and_validate()
    dev = testing.make_fake_dev(str(tmp_path), 1)
    proc = testing.make_fake_mig(str(tmp_path), dev, 1, 3)
    ngm = mgr.GPUManager(dev, proc, [], cfg, nvml=nvml.MockNvml(dev))
    with pytest.raises(RuntimeError, match=r"Number of partitions \(3\) for GPU 0 does not match expected partition count \(7\)"):
        ngm.start()


def test_serves_without_kubelet_socket(harness):
    h = harness(with_kubelet=False)
    c = h.connect()
    _, devs = first_list(c)
    assert set(devs) == {"nvidia0", "nvidia1"}


def test_plugin_socket_removed_restarts_server(harness):
    import os
    h = harness()
    h.kubelet.wait_registration()
    h.connect()
    os.unlink(os.path.join(h.plugin_dir, h.endpoint))
    reg = h.kubelet.wait_registration(timeout=10)
    assert reg.endpoint == h.endpoint


def test_transport_hook_exports_b200coll_profile(harness):
    h = harness(GPUConfig(transport=TransportConfig(name="b200coll", env={"B200COLL_ALGO": "nvls"})))
    c = h.connect()
    cr = c.allocate(["nvidia0"]).container_responses[0]
    env = dict(cr.envs)
    assert env["B200COLL_LIB"] == "/usr/local/nvidia/lib64/libb200coll.so" and env["LD_LIBRARY_PATH"] == "/usr/local/nvidia/lib64"
    assert env["B200COLL_ALGO"] == "nvls"
    assert len(cr.mounts) == 2      # lib dir already covered by the /usr/local/nvidia mount


# ------------------------------------------------------------------------------------------------- preferred allocation (opt-in)

def test_preferred_allocation_policy_function():
    numa = {f"nvidia{i}": (0 if i < 4 else 1) for i in range(8)}.get
    free = [f"nvidia{i}" for i in range(8)]
    # a 4-GPU request fits one socket: never straddle; natural order inside the node
    assert preferred.preferred_allocation(free, [], 4, numa) == ["nvidia0", "nvidia1", "nvidia2", "nvidia3"]
    # 2 free on socket 0, 4 on socket 1, want 3: socket 0 cannot cover it, socket 1 can -> all from socket 1
    assert preferred.preferred_allocation(["nvidia1", "nvidia3", "nvidia4", "nvidia5", "nvidia6", "nvidia7"], [], 3, numa) == ["nvidia4", "nvidia5", "nvidia6"]
    # must-include pins the socket; the rest follows it
    assert preferred.preferred_allocation(free, ["nvidia6"], 3, numa) == ["nvidia6", "nvidia4", "nvidia5"]
    # nothing fits one socket (want 6): take the fuller socket first, then spill
    got = preferred.preferred_allocation(["nvidia0", "nvidia1", "nvidia4", "nvidia5", "nvidia6", "nvidia7", "nvidia2"], [], 6, numa)
    assert got[:4] == ["nvidia4", "nvidia5", "nvidia6", "nvidia7"] and set(got[4:]) <= {"nvidia0", "nvidia1", "nvidia2"}
    assert preferred.preferred_allocation(["nvidia10", "nvidia2"], [], 1, lambda d: None) == ["nvidia2"]                 # natural, not lexicographic, order
    # shared GPUs: nvidia0 has 1 free replica (busy), nvidia1 has 3 (idle). spread -> the idle GPU; packed -> fill the busy one
    shared = ["nvidia0/vgpu2", "nvidia1/vgpu0", "nvidia1/vgpu1", "nvidia1/vgpu2"]
    assert preferred.preferred_allocation(shared, [], 1, lambda d: 0, "spread") == ["nvidia1/vgpu0"]
    assert preferred.preferred_allocation(shared, [], 1, lambda d: 0, "packed") == ["nvidia0/vgpu2"]
    # two replicas: packed keeps them on one physical GPU, spread takes two different ones
    assert [preferred.physical_of(d) for d in preferred.preferred_allocation(shared[1:], [], 2, lambda d: 0, "packed")] == ["nvidia1", "nvidia1"]
    assert len({preferred.physical_of(d) for d in preferred.preferred_allocation(shared, [], 2, lambda d: 0, "spread")}) == 2
    assert preferred.preferred_allocation(["nvidia0"], [], 3, lambda d: 0) == ["nvidia0"]                               # fewer free than wanted
    with pytest.raises(ValueError):
        preferred.preferred_allocation(free, [], 1, numa, "none")


def test_preferred_allocation_over_grpc_is_opt_in(harness):
    h = harness(gpus=4, policy="spread", numa=[0, 0, 1, 1])
    reg = h.kubelet.wait_registration()
    assert reg.options.get_preferred_allocation_available and not reg.options.pre_start_required
    c = h.connect()
    assert c.options().get_preferred_allocation_available
    _, devs = first_list(c)
    assert [n.ID for n in devs["nvidia2"].topology.nodes] == [1]
    assert c.preferred(["nvidia0", "nvidia2", "nvidia3"], [], 2) == ["nvidia2", "nvidia3"]          # the socket that can hold both
    assert c.preferred(["nvidia0", "nvidia1", "nvidia2", "nvidia3"], ["nvidia3"], 2) == ["nvidia3", "nvidia2"]


def test_without_the_flag_the_contract_is_the_references(harness):
    h = harness()
    reg = h.kubelet.wait_registration()
    assert not reg.HasField("options")
    c = h.connect()
    assert not c.options().get_preferred_allocation_available
    assert c.preferred(["nvidia0", "nvidia1"], [], 1) == []        # stub answer, as in the reference (logs an error)
