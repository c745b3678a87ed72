"""GPU-side tools (tools/*.cu) on real hardware. Runs last (file name) so a surprise here cannot hide the library's results.
The fault injector is opt-in: it does what its name says to the GPU it runs on."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "build")
pytestmark = pytest.mark.gpu


def run(*argv, env=None, timeout=120):
    return subprocess.run(list(argv), capture_output=True, text=True, timeout=timeout, env={**os.environ, **(env or {})})


def test_mps_probe_reports_memory_and_sm_count():
    """Role of reference example/cuda-mps/cuda_mem_and_sm_count.c:19-59, plus --json for assertions."""
    r = run(os.path.join(BUILD, "mps_probe"), "--json", env={"CUDA_MPS_ACTIVE_THREAD_PERCENTAGE": "25"})
    assert r.returncode == 0, r.stderr
    doc = json.loads(r.stdout)
    assert doc["active_thread_percentage"] == "25"          # echoed, so an e2e can tie the numbers to the limit it set
    dev = doc["devices"][0]
    assert dev["sm_count"] == 148 and dev["total_mib"] > 150_000 and 0 < dev["free_mib"] <= dev["total_mib"]      # no MPS daemon here: the full B200
    text = run(os.path.join(BUILD, "mps_probe"))
    assert "multiProcessorCount: 148" in text.stdout and "Free memory" in text.stdout


def test_nvls_probe_describes_the_box():
    r = run(os.path.join(BUILD, "nvls_probe"))
    # exit status depends on the box (a single GPU has no peer to multicast to); the description itself must always come out
    assert "driver_version" in r.stdout and "multicast=" in r.stdout and "sms=148" in r.stdout, r.stdout + r.stderr



@pytest.mark.parametrize("mode", ["oob-store", "oob-load", "trap"])
def test_xid_inject_faults_its_own_context_only(mode):
    """Reference demo/gpu-error/illegal-memory-access/vectorAdd.cu:28-70: the out-of-bounds store must be reported by the driver;
    the process exits 1 and the GPU stays usable for the next process."""
    r = run(os.path.join(BUILD, "xid_inject"), "--mode", mode)
    assert r.returncode == 1 and "fault raised as expected" in r.stderr, r.stdout + r.stderr
    after = run(os.path.join(BUILD, "mps_probe"), "--json")
    assert after.returncode == 0 and json.loads(after.stdout)["devices"][0]["sm_count"] == 148
