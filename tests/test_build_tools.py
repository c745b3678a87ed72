"""Lint tooling and the job generator (roles of reference build/boilerplate/boilerplate.py, build/check_*.sh, demo/gpu-training/generate_job.sh)."""
import os
import subprocess
import sys

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tree_passes_presubmit():
    assert subprocess.run([sys.executable, os.path.join(ROOT, "build_tools", "boilerplate.py")], capture_output=True, text=True).returncode == 0
    r = subprocess.run(["bash", os.path.join(ROOT, "build_tools", "check_style.sh")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_boilerplate_flags_missing_header(tmp_path):
    good = tmp_path / "good.py"; good.write_text('#!/usr/bin/env python\n"""Does a thing."""\nx = 1\n')
    bad = tmp_path / "bad.py"; bad.write_text("import os\n")
    badsh = tmp_path / "bad.sh"; badsh.write_text("#!/bin/bash\nset -e\n")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "build_tools", "boilerplate.py"), "--rootdir", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 1 and "bad.py" in r.stdout and "bad.sh" in r.stdout and "good.py" not in r.stdout


def test_job_generator_emits_32_valid_jobs(tmp_path):
    r = subprocess.run(["bash", os.path.join(ROOT, "demo", "gpu-training", "generate_job.sh")], env={**os.environ, "OUT_DIR": str(tmp_path / "jobs")}, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    files = sorted(os.listdir(tmp_path / "jobs"))
    assert len(files) == 32                                   # 4 depths x 2 batch sizes x 4 learning rates (reference generate_job.sh:30-81)
    job = yaml.safe_load(open(tmp_path / "jobs" / files[0]))
    c = job["spec"]["template"]["spec"]["containers"][0]
    assert job["kind"] == "Job" and c["resources"]["limits"]["nvidia.com/gpu"] == 8 and any(a.startswith("--resnet_size=") for a in c["args"])


def test_shipped_tuner_table_matches_builtin(coll_lib):
    """coll/tuner/b200_nvswitch.tbl must describe the same policy as the table compiled into the library."""
    code = ("from container_engine_accelerators_b200.ops import coll\n"
            "out=[]\n"
            "for op in range(4):\n"
            "  for n in (2,4,8):\n"
            "    for nvls in (0,1):\n"
            "      for b in (1024, 300<<10, 1<<20, 3<<20, 1<<28): out.append(coll.tuner_pick(op,b,n,bool(nvls)))\n"
            "print(' '.join(out))")
    base = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, env={k: v for k, v in os.environ.items() if k != "B200COLL_TUNER_FILE"})
    over = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, env={**os.environ, "B200COLL_TUNER_FILE": os.path.join(ROOT, "coll", "tuner", "b200_nvswitch.tbl")})
    assert base.returncode == 0 and over.returncode == 0, base.stderr + over.stderr
    assert base.stdout == over.stdout


def test_dockerfiles_only_copy_files_that_exist():
    import subprocess, sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "build_tools", "check_dockerfiles.py")], capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout


def test_every_first_party_image_in_the_manifests_has_a_build_rule():
    """deploy/generate.py names the images; the Makefile must build and push exactly those (same registry, name and tag)."""
    import subprocess, sys
    sys.path.insert(0, os.path.join(ROOT, "deploy"))
    import generate
    wanted = {v for v in generate.IMG.values() if v.startswith(generate.REG + "/")}
    out = subprocess.run(["make", "-n", "push"], cwd=ROOT, capture_output=True, text=True, check=True).stdout
    built = {w for line in out.splitlines() for w in line.split() if w.startswith(generate.REG + "/")}
    pushed = out.strip().splitlines()[-1]
    assert wanted <= built, sorted(wanted - built)
    for image in wanted:
        name = image.split("/")[-1].split(":")[0]
        assert f" {name} " in pushed.replace(";", " ; ") or f" {name};" in pushed, (name, pushed)


def test_clock_sampler_without_nvml_and_with_the_nvidia_smi_fallback(tmp_path, monkeypatch):
    """bench.py's clocks block: no NVML and no nvidia-smi -> an empty summary, never an exception; with only nvidia-smi available the
    recipe's CSV is parsed (median of the upper half of the SM clock samples, throttle reasons collected)."""
    import sys
    import time
    from container_engine_accelerators_b200.utils.clocks import ClockSampler
    monkeypatch.setitem(sys.modules, "pynvml", None)                     # import pynvml -> ImportError
    monkeypatch.setenv("PATH", str(tmp_path))
    with ClockSampler(0) as c:
        pass
    assert c.summary() == {"sm_mhz": None, "sm_max_mhz": None, "power_w_max": None, "reasons": [], "samples": 0, "source": "none"}
    smi = tmp_path / "nvidia-smi"
    smi.write_text("#!/bin/bash\nfor mhz in 1200 1965 1965 1950; do echo \"0, $mhz, 1965, 700.5, 0x4, Not Active, Not Active, Not Active, Active\"; done\nsleep 5\n")
    smi.chmod(0o755)
    monkeypatch.setenv("PATH", f"{tmp_path}:/usr/bin:/bin")
    with ClockSampler(0) as c:
        time.sleep(0.3)
    s = c.summary()
    assert s["source"] == "nvidia-smi" and s["samples"] == 4 and s["sm_mhz"] == 1965.0 and s["sm_max_mhz"] == 1965.0 and s["power_w_max"] == 700.5 and s["reasons"] == ["sw_power_cap"]
