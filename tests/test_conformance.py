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
    assert r.stdout.count("PASS") == 9


def test_go_style_flags_are_accepted_by_the_python_agent():
    from container_engine_accelerators_b200.agent import main as agent_main
    argv = agent_main.go_style_argv(["-enable-health-monitoring", "-gpu-config=/x/y.json", "-v", "3", "--host-path", "/h", "-logtostderr"])
    assert argv == ["--enable-health-monitoring", "--gpu-config=/x/y.json", "-v", "3", "--host-path", "/h", "--logtostderr"]
    args = agent_main.build_parser().parse_args(argv)
    assert args.enable_health_monitoring and args.gpu_config == "/x/y.json" and args.verbosity == 3 and args.host_path == "/h"
