"""The node agent against a REAL B200: the native NVML binding on the installed libnvidia-ml, both device-plugin implementations
serving the box's own /dev/nvidia* to a stub kubelet, and the Xid health loop fed by a real fault (the injector raises an MMU fault,
the driver logs an Xid, NVML delivers the event, the plugin flips the device to Unhealthy). Everything else about the agent is tested
against scripted fakes on CPU (tests/test_device_plugin.py, test_native_device_plugin.py, conformance/run.py); these tests are the
once-on-hardware counterpart the fakes cannot give (reference: pkg/gpu/nvidia/beta_plugin_test.go:36-70,343-386 run the Go plugin
against a KubeletStub; health_check/health_checker.go:452-468 registers for Xid events on real NVML;
demo/gpu-error/illegal-memory-access/vectorAdd.cu:28-34 is the fault).
The file sorts last on purpose: the fault injection kills CUDA contexts (its own), so it runs after the collective tests."""
import os
import re
import subprocess
import sys
import time

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BUILD = os.path.join(ROOT, "build")


def _minors():
    return sorted(int(m.group(1)) for m in (re.fullmatch(r"nvidia(\d+)", f) for f in os.listdir("/dev")) if m)


@pytest.fixture(scope="module")
def gpu_count():
    import torch
    if not torch.cuda.is_available() or not _minors():
        pytest.skip("no CUDA device / no /dev/nvidia<N>")
    return torch.cuda.device_count()


def _first_gpu() -> str:
    """The plugin names devices after their /dev node; a one-GPU lease of an eight-GPU host may expose /dev/nvidia5 only."""
    return f"nvidia{_minors()[0]}"


def _busy(seconds: float) -> None:
    import torch
    a = torch.randn(4096, 4096, device="cuda", dtype=torch.bfloat16)
    t0 = time.time()
    while time.time() - t0 < seconds:
        for _ in range(20):
            a = (a @ a).clamp_(-1, 1)
        torch.cuda.synchronize()


def test_native_nvml_binding_on_the_installed_driver(gpu_count):
    """libb200agent_nvml.so dlopens the real libnvidia-ml.so.1: enumeration, the utilisation sampler (zero-sample guard included) and
    Xid event registration (reference: pkg/gpu/nvidia/metrics/util.go:37-87, health_check/health_checker.go:452-468)."""
    from container_engine_accelerators_b200.agent import nvml
    env_backup = os.environ.pop("B200AGENT_NVML_LIB", None)
    try:
        n = nvml.NativeNvml()
        n.init()
        assert n.device_count() >= gpu_count >= 1
        minors = _minors()
        d = next((x for x in (n.device(i) for i in range(n.device_count())) if x.minor in minors), None)     # NVML may list GPUs this container has no node for
        assert d is not None, "no NVML device matches a /dev/nvidia<N> node"
        assert "B200" in d.name and d.uuid.startswith("GPU-") and re.fullmatch(r"[0-9A-Fa-f]{8}:[0-9A-Fa-f]{2}:[0-9A-Fa-f]{2}\.[0-9]", d.bus_id), d
        assert 170 << 30 < d.mem_total < 200 << 30 and d.minor >= 0 and os.path.exists(f"/dev/nvidia{d.minor}")
        assert re.match(r"\d+\.\d+", n.driver_version())
        assert nvml.numa_topology(d.bus_id) in (None, 0, 1, 2, 3)
        try:
            idle = n.average_usage(d.uuid, int((time.time() - 1) * 1e6))
        except nvml.NvmlError:
            idle = 0                                            # an idle GPU may have produced no sample in the last second: reported as an error, not a crash
        assert 0 <= idle <= 100
        _busy(1.5)
        busy = n.average_usage(d.uuid, int((time.time() - 1.0) * 1e6))
        assert 0 < busy <= 100, busy                            # the sampler saw the matmuls; never above 100 (the reference's cgo helper could divide by zero)
        with pytest.raises(nvml.NvmlError):                    # a window with no samples is an error code (the reference's cgo helper divided by zero here)
            n.average_usage(d.uuid, int((time.time() + 3600) * 1e6))
        h = n.events_open()
        assert n.events_register(h, d.index) is True            # a B200 supports Xid events
        assert n.events_wait(h, 200) is None                    # nothing is wrong: timeout, not an error
        n.events_close(h)
        n.shutdown()
    finally:
        if env_backup is not None:
            os.environ["B200AGENT_NVML_LIB"] = env_backup


@pytest.mark.parametrize("impl", ["native", "python"])
def test_device_plugin_serves_the_real_gpus_to_a_stub_kubelet(gpu_count, impl):
    from conformance import run as conf
    n = conf.Node(conf.PRESETS[impl], real=True)
    try:
        reg = n.kubelet.wait_registration(30)
        assert (reg.version, reg.resource_name) == ("v1beta1", "nvidia.com/gpu")
        c = n.connect()
        stream, devs = conf.first_list(c)
        assert set(devs) == {f"nvidia{i}" for i in _minors()} and all(d.health == "Healthy" for d in devs.values()), devs
        gpu = _first_gpu()
        topo = [x.ID for x in devs[gpu].topology.nodes]
        assert topo in ([], [0], [1], [2], [3])                 # the NUMA node sysfs reports for the GPU's PCI function (absent on single-node hosts)
        cr = c.allocate([gpu]).container_responses[0]
        paths = [d.host_path for d in cr.devices]
        assert paths[0] == f"/dev/{gpu}" and "/dev/nvidiactl" in paths and "/dev/nvidia-uvm" in paths
        assert all(os.path.exists(p) for p in paths), paths      # only device nodes that exist on this host are handed to the container
        assert all(d.permissions == "mrw" for d in cr.devices) and len(cr.mounts) == 2
        conf.expect_error(lambda: c.allocate(["nvidia99"]), "non-existing device nvidia99")
        stream.cancel()
    finally:
        n.close()


def test_metrics_sampler_reports_the_real_gpu(gpu_count):
    """The native plugin's /metrics endpoint on real NVML: per-node duty cycle and memory gauges for GPU 0 carry its real UUID and model."""
    import socket
    import urllib.request
    from conformance import run as conf
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    n = conf.Node(conf.PRESETS["native"], real=True, extra=f"-enable-container-gpu-metrics -gpu-metrics-port {port} -gpu-metrics-collection-interval 500")
    try:
        n.connect()
        _busy(1.0)
        deadline, body = time.time() + 20, ""
        while time.time() < deadline:
            try:
                body = urllib.request.urlopen(f"http://127.0.0.1:{port}/metrics", timeout=2).read().decode()
                if "duty_cycle_gpu_node" in body and "B200" in body:
                    break
            except OSError:
                pass
            time.sleep(0.5)
        assert "duty_cycle_gpu_node" in body and "memory_total_gpu_node" in body and 'model="NVIDIA B200"' in body, body[:2000] + n.logs()[-2000:]
        m = re.search(r'memory_total_gpu_node\{[^}]*\} ([0-9.e+]+)', body)
        assert m and 170e9 < float(m.group(1)) < 200e9 * 1.1
    finally:
        n.close()


def test_real_xid_marks_the_device_unhealthy(gpu_count):
    """End to end on hardware: the plugin's health loop waits on a real NVML event set; `xid_inject` makes an out-of-bounds store from a
    kernel (an MMU fault: Xid 31 on this driver; 13 / 43 on others); the ListAndWatch stream must then report the GPU Unhealthy and
    Allocate must refuse it (reference: health_check/health_checker.go:355-449, demo/gpu-error/illegal-memory-access)."""
    from conformance import run as conf
    inject = os.path.join(BUILD, "xid_inject")
    if not os.path.exists(inject):
        pytest.skip("build/xid_inject missing (make -C tools)")
    n = conf.Node(conf.PRESETS["native"], real=True, extra="-enable-health-monitoring", env={"XID_CONFIG": "13,31,43,45,109"})
    try:
        c = n.connect()
        stream, devs = conf.first_list(c)
        gpu = _first_gpu()
        assert devs[gpu].health == "Healthy"
        time.sleep(2.0)                                           # the health loop registers for events after the first list is out
        r = subprocess.run([inject, "--mode", "oob-store"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 1 and "fault raised as expected" in r.stderr, r.stdout + r.stderr
        import threading
        got = {}

        def watch():
            try:
                for resp in stream:
                    health = {d.ID: d.health for d in resp.devices}
                    if health.get(gpu) == "Unhealthy":
                        got["health"] = health
                        return
            except Exception as e:      # stream cancelled at the end of the test
                got.setdefault("error", repr(e))
        t = threading.Thread(target=watch, daemon=True)
        t.start()
        t.join(30)
        if "health" not in got:
            log = n.logs()
            if "Xid" not in log and "xid" not in log:
                pytest.skip("this container does not receive NVML Xid events for the fault (no event in 30 s); plugin log:\n" + log[-1500:])
            raise AssertionError(f"the plugin saw an Xid but never reported {gpu} Unhealthy:\n" + log[-3000:])
        assert got["health"][gpu] == "Unhealthy"
        conf.expect_error(lambda: c.allocate([gpu]), f"unhealthy device {gpu}")
        stream.cancel()
    finally:
        n.close()
    after = subprocess.run([os.path.join(BUILD, "mps_probe"), "--json"], capture_output=True, text=True, timeout=60)
    assert after.returncode == 0, after.stdout + after.stderr     # the fault killed the injector's context only
