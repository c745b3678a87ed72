"""conformance/run.py against both implementations, as processes (the Python agent's `main()` boot path included)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("impl", ["native", "python"])
def test_conformance_suite(impl, native_build):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "conformance", "run.py"), "--impl", impl], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK: all scenarios passed" in r.stdout, r.stdout + r.stderr
    assert r.stdout.count("PASS") == 15


def test_go_style_flags_are_accepted_by_the_python_agent():
    from container_engine_accelerators_b200.agent import main as agent_main
    argv = agent_main.go_style_argv(["-enable-health-monitoring", "-gpu-config=/x/y.json", "-v", "3", "--host-path", "/h", "-logtostderr"])
    assert argv == ["--enable-health-monitoring", "--gpu-config=/x/y.json", "-v", "3", "--host-path", "/h", "--logtostderr"]
    args = agent_main.build_parser().parse_args(argv)
    assert args.enable_health_monitoring and args.gpu_config == "/x/y.json" and args.verbosity == 3 and args.host_path == "/h"


def test_python_agent_process_with_kubernetes_side_effects(tmp_path, native_build):
    """The Python agent's main() with health monitoring, driver-version publishing and the API server reached through B200_KUBE_URL:
    Xid event -> Unhealthy over ListAndWatch, Event + Node condition + annotations in the (fake) API server; then --status-only mode,
    which does the Kubernetes side only and serves no plugin socket."""
    import signal
    import time

    import yaml

    from container_engine_accelerators_b200.agent import testing
    from container_engine_accelerators_b200.agent.plugin import DevicePluginClient
    api = testing.FakeKubeApi().start()
    procs = []
    try:
        api.add_node("node-a", boot_id="boot-1")
        dev = testing.make_fake_dev(str(tmp_path), 2)
        plugin_dir = tmp_path / "dp"; plugin_dir.mkdir()
        events = tmp_path / "events.txt"; events.write_text("")
        env = {**os.environ, "PYTHONPATH": ROOT, "B200AGENT_NVML_LIB": os.path.join(native_build, "libfake_nvml.so"), "B200AGENT_NATIVE_LIB": os.path.join(native_build, "libb200agent_nvml.so"),
               "FAKE_NVML_DEV_DIR": dev, "FAKE_NVML_EVENTS": str(events), "FAKE_NVML_DRIVER": "570.124.06", "B200_KUBE_URL": api.url, "NODE_NAME": "node-a", "XID_CONFIG": "31"}
        base = [sys.executable, "-m", "container_engine_accelerators_b200.agent.main", "-plugin-directory", str(plugin_dir), "-gpu-config", str(tmp_path / "none.json"),
                "--dev-directory", dev, "--proc-directory", str(tmp_path / "proc"), "-enable-health-monitoring", "-publish-driver-version", "--socket-check-interval", "0.1"]
        log = open(tmp_path / "agent.log", "w")
        p = subprocess.Popen(base + ["--plugin-endpoint", "nvidiaGPU-py.sock"], env=env, stdout=log, stderr=log); procs.append(p)
        sock = plugin_dir / "nvidiaGPU-py.sock"
        deadline = time.time() + 30
        while not sock.exists():
            assert time.time() < deadline and p.poll() is None, open(log.name).read()
            time.sleep(0.05)
        c = DevicePluginClient(str(sock)); c.wait_ready()
        stream = c.list_and_watch(); next(stream)
        time.sleep(1.0)
        with open(events, "a") as f:
            f.write("1 48\n")
        assert {d.ID: d.health for d in next(stream).devices} == {"nvidia0": "Healthy", "nvidia1": "Unhealthy"}
        while time.time() < deadline and not (api.events and any(c2["type"] == "XidCriticalError" for c2 in api.nodes["node-a"]["status"]["conditions"])
                                              and "cloud.google.com/cuda.driver-version.full" in api.nodes["node-a"]["metadata"]["annotations"]):
            time.sleep(0.05)
        assert api.events[0]["message"] == "Caught XID error, XID=48"
        assert api.nodes["node-a"]["metadata"]["annotations"]["cloud.google.com/cuda.driver-version.major"] == "570"
        stream.cancel(); c.close()
        p.send_signal(signal.SIGTERM)
        assert p.wait(10) == 0 and not sock.exists()                      # graceful: exit 0, plugin socket removed
        # status-only: no plugin socket, but the condition heartbeat / events path still runs
        api.events.clear()
        events2 = tmp_path / "events2.txt"; events2.write_text("")          # the scripted NVML replays its file from the start in a new process
        p2 = subprocess.Popen(base + ["--status-only", "--plugin-endpoint", "nvidiaGPU-status.sock"], env={**env, "FAKE_NVML_EVENTS": str(events2)}, stdout=log, stderr=log); procs.append(p2)
        time.sleep(2.0)
        with open(events2, "a") as f:
            f.write("0 79\n")
        while time.time() < deadline + 20 and not api.events:
            time.sleep(0.05)
        assert api.events and api.events[0]["message"] == "Caught XID error, XID=79"
        assert not (plugin_dir / "nvidiaGPU-status.sock").exists()
        role = next(d for d in yaml.safe_load_all(open(os.path.join(ROOT, "deploy", "device-plugin", "rbac.yaml"))) if d["kind"] == "ClusterRole")
        assert testing.rbac_violations(role, api.requests) == []
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
        api.stop()
