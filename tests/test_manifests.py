"""deploy/: the generated tree is up to date, parses, and keeps the invariants the node agents rely on; the embedded shell
programs run against stubs. The reference has no tests for any manifest or script."""
import glob
import os
import subprocess
import sys

import pytest
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEPLOY = os.path.join(ROOT, "deploy")
SCRIPTS = os.path.join(DEPLOY, "scripts")


def docs(rel):
    with open(os.path.join(DEPLOY, rel)) as f:
        return [d for d in yaml.safe_load_all(f) if d]


def test_generated_tree_is_current():
    r = subprocess.run([sys.executable, os.path.join(DEPLOY, "generate.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_every_manifest_parses_and_has_kind():
    files = glob.glob(os.path.join(DEPLOY, "**", "*.yaml"), recursive=True)
    assert len(files) >= 74                      # the reference ships 74 YAML files
    for f in files:
        for d in yaml.safe_load_all(open(f)):
            assert d and "kind" in d and "apiVersion" in d, f


def test_device_plugin_daemonset_contract():
    (ds,) = docs("device-plugin/device-plugin.yaml")
    spec = ds["spec"]["template"]["spec"]
    c = spec["containers"][0]
    assert "--enable-container-gpu-metrics" in c["command"] and "--enable-health-monitoring" in c["command"]
    paths = {v["name"]: v["hostPath"]["path"] for v in spec["volumes"]}
    assert paths["device-plugin"] == "/var/lib/kubelet/device-plugins" and paths["nvidia"] == "/home/kubernetes/bin/nvidia" and paths["pod-resources"] == "/var/lib/kubelet/pod-resources"
    mounts = {m["name"]: m["mountPath"] for m in c["volumeMounts"]}
    assert mounts["device-plugin"] == "/device-plugin" and mounts["nvidia"] == "/usr/local/nvidia" and mounts["nvidia-config"] == "/etc/nvidia"
    env = {e["name"]: e for e in c["env"]}
    assert env["XID_CONFIG"]["valueFrom"]["configMapKeyRef"] == {"name": "xid-config", "key": "HealthCriticalXid", "optional": True}
    assert env["NODE_NAME"]["valueFrom"]["fieldRef"]["fieldPath"] == "spec.nodeName"          # not GCE metadata
    assert c["securityContext"]["privileged"] and spec["priorityClassName"] == "system-node-critical"
    assert c["ports"][0]["containerPort"] == 2112
    rbac = docs("device-plugin/rbac.yaml")
    role = [d for d in rbac if d["kind"] == "ClusterRole"][0]
    assert any("nodes/status" in r["resources"] and "patch" in r["verbs"] for r in role["rules"])


def test_installer_shape_init_does_the_work_pause_keeps_running():
    for rel in ("transport/b200coll-installer.yaml", "transport/compat/nccl-rdma-installer.yaml", "driver-installer/cos/daemonset-preloaded-latest.yaml", "partition-gpu/partition-gpu.yaml"):
        ds = [d for d in docs(rel) if d["kind"] == "DaemonSet"][0]
        spec = ds["spec"]["template"]["spec"]
        assert spec["initContainers"] and spec["containers"][-1]["name"] == "pause", rel
        assert spec["tolerations"] == [{"operator": "Exists"}] and spec["hostNetwork"] and spec["hostPID"], rel
        assert ds["spec"]["updateStrategy"]["type"] == "RollingUpdate"


def test_b200coll_installer_targets_b200_and_mounts_host_lib_dir():
    (ds,) = docs("transport/b200coll-installer.yaml")
    spec = ds["spec"]["template"]["spec"]
    terms = spec["affinity"]["nodeAffinity"]["requiredDuringSchedulingIgnoredDuringExecution"]["nodeSelectorTerms"][0]["matchExpressions"]
    assert {"key": "cloud.google.com/gke-accelerator", "operator": "In", "values": ["nvidia-b200"]} in terms
    init = spec["initContainers"][0]
    assert "libb200coll.so" in init["command"][2] and "--selfcheck" in init["command"][2]
    assert {"name": "library-dir-host", "mountPath": "/usr/local/nvidia"} in init["volumeMounts"]


def test_mig_and_a4x_variants():
    mig = docs("driver-installer/cos/daemonset-nvidia-mig.yaml")[0]["spec"]["template"]["spec"]
    assert [c["name"] for c in mig["initContainers"]] == ["nvidia-driver-installer", "partition-gpus"]
    a4x = docs("driver-installer/cos/daemonset-preloaded-latest-a4x.yaml")[0]["spec"]["template"]["spec"]
    assert [c["name"] for c in a4x["initContainers"]] == ["nvidia-driver-installer"]                     # GB200: no MIG step
    conf = docs("driver-installer/cos/daemonset-confidential-latest.yaml")[0]["spec"]["template"]["spec"]
    assert conf["initContainers"][1]["restartPolicy"] == "Always" and "confidential_node_type.txt" in conf["initContainers"][0]["command"][2]
    vgpu = docs("driver-installer/cos/daemonset-vgpu-latest.yaml")[0]["spec"]["template"]["spec"]
    assert vgpu["containers"][0]["name"] == "nvidia-daemon-installer" and "machine_type.txt" in vgpu["initContainers"][0]["command"][2]
    for rel, ver in (("R525", "525"), ("R570", "570")):
        env = {e["name"]: e.get("value") for e in docs(f"driver-installer/ubuntu/daemonset-preloaded-{rel}.yaml")[0]["spec"]["template"]["spec"]["initContainers"][0]["env"]}
        assert env["NVIDIA_DRIVER_VERSION"].startswith(ver)


def test_nccl_test_pods():
    a4 = [d for d in docs("nccl-test/rdma/nccl-test-a4.yaml") if d["kind"] == "Pod"]
    assert len(a4) == 2
    import json
    ifaces = json.loads(a4[0]["metadata"]["annotations"]["networking.gke.io/interfaces"])
    assert len(ifaces) == 9 and ifaces[1] == {"interfaceName": "eth2", "network": "rdma-0"}               # 8 RDMA NICs
    c = a4[0]["spec"]["containers"][0]
    assert c["resources"]["limits"]["nvidia.com/gpu"] == 8 and "set_nccl_env.sh" in c["args"][0]
    shm = [v for v in a4[0]["spec"]["volumes"] if v["name"] == "shared-memory"][0]
    assert shm["emptyDir"] == {"medium": "Memory", "sizeLimit": "250Gi"}
    tcpxo = [d for d in docs("nccl-test/tcpxo/nccl-test-latest.yaml") if d["kind"] == "Pod"][0]
    ann = tcpxo["metadata"]["annotations"]
    injected = yaml.safe_load(ann["devices.gke.io/container.tcpxo-daemon"])
    assert {"path": "/dev/dmabuf_import_helper"} in injected and len(injected) == 11                      # 8 GPUs + nvidiactl + nvidia-uvm + helper
    assert tcpxo["spec"]["initContainers"][0]["restartPolicy"] == "Always"
    imex = docs("nccl-test/rdma/nccl-test-imex-a4x.yaml")
    assert imex[0]["kind"] == "ComputeDomain" and [d for d in imex if d["kind"] == "Pod"][0]["spec"]["containers"][0]["resources"]["limits"]["nvidia.com/gpu"] == 4
    js = [d for d in docs("nccl-test/rdma/nccl-test-a4x-max-jobset.yaml") if d["kind"] == "JobSet"][0]
    assert "all_gather_perf" in str(js) and "-b 1K -e 8G -f 2 -g 1 -w 5 --iters 100 -c 1" in str(js)
    ours = docs("nccl-test/b200coll-test.yaml")[0]
    script = ours["spec"]["containers"][0]["args"][0]
    assert "${op}_perf --procs --ranks 8" in script and "broadcast reduce" in script          # nccl-tests names, all six collectives


def test_scripts_have_valid_syntax():
    for f in glob.glob(os.path.join(SCRIPTS, "*.sh")):
        assert subprocess.run(["bash", "-n", f]).returncode == 0, f


def stub(tmp_path, name, body="exit 0"):
    p = tmp_path / name
    p.write_text(f"#!/bin/bash\necho \"{name} $*\" >> {tmp_path}/calls.log\n{body}\n")
    p.chmod(0o755)
    return str(p)


def test_cos_driver_install_skips_when_module_loaded(tmp_path):
    inst = stub(tmp_path, "cos-gpu-installer")
    (tmp_path / "root/home/kubernetes/bin/nvidia").mkdir(parents=True)
    env = {**os.environ, "COS_GPU_INSTALLER": inst, "ROOT_MOUNT_DIR": str(tmp_path / "root"), "LSMOD": stub(tmp_path, "lsmod", "echo 'nvidia 123 0'")}
    assert subprocess.run(["bash", os.path.join(SCRIPTS, "cos-driver-install.sh")], env=env).returncode == 0
    assert "cos-gpu-installer" not in (tmp_path / "calls.log").read_text()
    env["LSMOD"] = stub(tmp_path, "lsmod2", "echo 'ext4 1 1'")
    assert subprocess.run(["bash", os.path.join(SCRIPTS, "cos-driver-install.sh")], env=env).returncode == 0
    assert "cos-gpu-installer install --version=latest" in (tmp_path / "calls.log").read_text()
    assert oct((tmp_path / "root/home/kubernetes/bin/nvidia").stat().st_mode)[-3:] == "755"


def test_transport_install_copies_tree(tmp_path):
    src = tmp_path / "src"; (src / "sub").mkdir(parents=True); (src / "libnccl-net.so").write_text("x"); (src / "sub" / "tuner.textproto").write_text("y")
    extra = tmp_path / "gib"; extra.mkdir(); (extra / "set_nccl_env.sh").write_text("z")
    env = {**os.environ, "TRANSPORT_SRC_DIR": str(src), "NCCL_INSTALL_DIR": str(tmp_path / "dst"), "TRANSPORT_ENTRY": stub(tmp_path, "container_entry.sh"),
           "TRANSPORT_EXTRA_SRC": str(extra), "TRANSPORT_EXTRA_DST": str(tmp_path / "gib-host")}
    assert subprocess.run(["bash", os.path.join(SCRIPTS, "transport-install.sh")], env=env).returncode == 0
    assert (tmp_path / "dst/libnccl-net.so").exists() and (tmp_path / "dst/sub/tuner.textproto").exists() and (tmp_path / "gib-host/set_nccl_env.sh").exists()
    assert "container_entry.sh install --install-nccl" in (tmp_path / "calls.log").read_text()


def test_b200coll_install_and_env_profile(tmp_path):
    src = tmp_path / "opt"; (src / "lib").mkdir(parents=True); (src / "bin").mkdir(); (src / "tuner").mkdir()
    for f in ("lib/libb200coll.so", "lib/libb200coll_nccl.so", "b200coll-env-profile.sh", "tuner/b200_nvswitch.tbl"):
        (src / f).write_text("x")
    (src / "bin").mkdir(exist_ok=True); (src / "bin" / "mps_probe").write_text("x")
    perf = src / "bin" / "b200coll_perf"; perf.write_text("#!/bin/bash\necho selfcheck-ran > %s/selfcheck\nexit 0\n" % tmp_path); perf.chmod(0o755)
    env = {**os.environ, "B200COLL_SRC_DIR": str(src), "NCCL_INSTALL_DIR": str(tmp_path / "lib64"), "B200COLL_BIN_DIR": str(tmp_path / "bin")}
    assert subprocess.run(["bash", os.path.join(SCRIPTS, "b200coll-install.sh")], env=env).returncode == 0
    assert (tmp_path / "lib64/libb200coll.so").exists() and (tmp_path / "lib64/b200_nvswitch.tbl").exists() and (tmp_path / "selfcheck").exists()
    assert (tmp_path / "bin" / "mps_probe").exists()
    for name in ("all_reduce_perf", "all_gather_perf", "reduce_scatter_perf", "alltoall_perf", "broadcast_perf", "reduce_perf", "sendrecv_perf", "gather_perf", "scatter_perf", "hypercube_perf"):       # nccl-tests names
        assert os.readlink(tmp_path / "bin" / name) == "b200coll_perf"
    out = subprocess.run(["bash", "-c", f"B200COLL_LIB_DIR={tmp_path}/lib64 source {SCRIPTS}/b200coll-env-profile.sh; echo $B200COLL_LIB $B200COLL_ALGO $B200COLL_TUNER_FILE"], capture_output=True, text=True).stdout.split()
    assert out == [f"{tmp_path}/lib64/libb200coll.so", "auto", f"{tmp_path}/lib64/b200_nvswitch.tbl"]


def test_tcpxo_host_prep_exposes_only_matching_pci_functions(tmp_path):
    pci = tmp_path / "pci"
    for bdf, ven, dev in (("0000:06:00.0", "0x1ae0", "0x0084"), ("0000:07:00.0", "0x1ae0", "0x0042"), ("0000:08:00.0", "0x10de", "0x2901")):
        d = pci / bdf; d.mkdir(parents=True); (d / "vendor").write_text(ven + "\n"); (d / "device").write_text(dev + "\n"); (d / "resource0").write_text("")
    env = {**os.environ, "SYS_PCI_DEVICES": str(pci), "APERTURE_DIR": str(tmp_path / "ap"), "IPTABLES": stub(tmp_path, "iptables"), "MODPROBE": stub(tmp_path, "modprobe"), "MOUNT": stub(tmp_path, "mount")}
    assert subprocess.run(["bash", os.path.join(SCRIPTS, "tcpxo-host-prep.sh")], env=env).returncode == 0
    log = (tmp_path / "calls.log").read_text()
    assert "modprobe import-helper" in log and "iptables -I INPUT -p tcp -m tcp -j ACCEPT" in log
    assert log.count("mount --bind") == 1 and "0000:06:00.0" in log and os.path.isdir(tmp_path / "ap/0000:06:00.0")


def test_optmem_and_confidential_scripts(tmp_path):
    (tmp_path / "net/core").mkdir(parents=True); (tmp_path / "net/core/optmem_max").write_text("20480\n")
    assert subprocess.run(["bash", os.path.join(SCRIPTS, "optmem-max.sh")], env={**os.environ, "PROC_SYS": str(tmp_path)}).returncode == 0
    assert (tmp_path / "net/core/optmem_max").read_text().strip() == "131072"
    root = tmp_path / "root"; (root / "home/kubernetes/bin/nvidia").mkdir(parents=True)
    nv = tmp_path / "nv"; (nv / "bin").mkdir(parents=True)
    (nv / "bin/nvidia-modprobe").write_text("#!/bin/bash\nexit 0\n"); (nv / "bin/nvidia-modprobe").chmod(0o755)
    env = {**os.environ, "ROOT_MOUNT_DIR": str(root), "NVIDIA_INSTALL_DIR_CONTAINER": str(nv), "COS_GPU_INSTALLER": stub(tmp_path, "installer"), "LSMOD": stub(tmp_path, "lsmod", "true"),
           "INSMOD": stub(tmp_path, "insmod"), "CURL": stub(tmp_path, "curl", "echo 'a=b,cloud.google.com/gke-confidential-nodes-instance-type=TDX,c=d'")}
    assert subprocess.run(["bash", os.path.join(SCRIPTS, "cos-confidential-install.sh")], env=env).returncode == 0
    assert (root / "etc/nvidia/confidential_node_type.txt").read_text().strip() == "TDX"
    log = (tmp_path / "calls.log").read_text()
    assert "installer install --version=latest --no-verify" in log and log.count("insmod ") == 4


# ------------------------------------------------------------------------------------------------- .run driver installers
DRV = os.path.join(DEPLOY, "driver-installer")


def run_installer(tmp_path, entry, extra_env=None, prime_cache=None):
    inst = tmp_path / "nvidia"; (inst / "bin").mkdir(parents=True, exist_ok=True)
    for tool in ("nvidia-smi", "nvidia-modprobe"):
        p = inst / "bin" / tool; p.write_text(f"#!/bin/bash\necho \"{tool} $*\" >> {tmp_path}/calls.log\n"); p.chmod(0o755)
    root = tmp_path / "root"; (root / "etc").mkdir(parents=True, exist_ok=True)
    (tmp_path / "ldconf").mkdir(exist_ok=True)
    if prime_cache:
        (inst / ".cache").write_text(prime_cache)
    env = {**os.environ, "NVIDIA_INSTALL_DIR_CONTAINER": str(inst), "NVIDIA_INSTALL_DIR_HOST": "/home/kubernetes/bin/nvidia", "ROOT_MOUNT_DIR": str(root), "KERNEL_VERSION": "6.8.0-1021-gke",
           "NVIDIA_DRIVER_VERSION": "570.124.06", "LD_SO_CONF_D": str(tmp_path / "ldconf"), "OVERLAY_ROOT": str(tmp_path / "ovl"),
           "MOUNT": stub(tmp_path, "mount"), "UMOUNT": stub(tmp_path, "umount"), "LDCONFIG": stub(tmp_path, "ldconfig"), "LSMOD": stub(tmp_path, "lsmod", "echo none"),
           "INSMOD": stub(tmp_path, "insmod"), "CURL": stub(tmp_path, "curl"), "SH": stub(tmp_path, "sh"), "APT_GET": stub(tmp_path, "apt-get"),
           "TAR": stub(tmp_path, "tar"), "MAKE": stub(tmp_path, "make"), "KERNEL_SRC_DIR": str(tmp_path / "ksrc"), **(extra_env or {})}
    (tmp_path / "ksrc/include/generated").mkdir(parents=True, exist_ok=True)
    r = subprocess.run(["bash", os.path.join(DRV, entry)], env=env, capture_output=True, text=True)
    log = (tmp_path / "calls.log").read_text() if (tmp_path / "calls.log").exists() else ""
    return r, log, inst, root


def test_ubuntu_installer_fresh_install(tmp_path):
    r, log, inst, root = run_installer(tmp_path, "ubuntu/entrypoint.sh")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "apt-get install -y linux-headers-6.8.0-1021-gke" in log
    assert log.count("mount -t overlay") == 3 and "lowerdir=/lib/modules/6.8.0-1021-gke/video" in log
    assert "NVIDIA-Linux-x86_64-570.124.06.run" in log and "--kernel-module-type=open" in log and "--no-drm --silent --accept-license" in log
    assert (inst / ".cache").read_text() == "CACHE_KERNEL_VERSION=6.8.0-1021-gke\nCACHE_NVIDIA_DRIVER_VERSION=570.124.06\n"
    assert "nvidia-modprobe -c0 -u" in log and log.count("umount") == 3
    assert (root / "etc/ld.so.conf").read_text() == "/home/kubernetes/bin/nvidia/lib64\n"


def test_ubuntu_installer_cache_hit_skips_download(tmp_path):
    r, log, inst, root = run_installer(tmp_path, "ubuntu/entrypoint.sh", prime_cache="CACHE_KERNEL_VERSION=6.8.0-1021-gke\nCACHE_NVIDIA_DRIVER_VERSION=570.124.06\n")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "curl" not in log and "apt-get" not in log and "mount -t overlay" not in log
    assert "insmod " + str(inst / "drivers/nvidia.ko") in log and "nvidia-uvm.ko" in log
    r2, _, _, root = run_installer(tmp_path, "ubuntu/entrypoint.sh", prime_cache="CACHE_KERNEL_VERSION=6.8.0-1021-gke\nCACHE_NVIDIA_DRIVER_VERSION=570.124.06\n")
    assert (root / "etc/ld.so.conf").read_text().count("nvidia/lib64") == 1          # idempotent


def test_old_driver_keeps_proprietary_modules_and_failure_propagates(tmp_path):
    r, log, _, _ = run_installer(tmp_path, "ubuntu/entrypoint.sh", extra_env={"NVIDIA_DRIVER_VERSION": "535.230.02"})
    assert r.returncode == 0 and "--kernel-module-type" not in log
    bad = stub(tmp_path, "sh-fail", "exit 3")
    r, log, inst, _ = run_installer(tmp_path / "x" if False else tmp_path, "ubuntu/entrypoint.sh", extra_env={"SH": bad, "NVIDIA_DRIVER_VERSION": "580.1.1"})
    assert r.returncode != 0


def test_minikube_installer_builds_kernel_source(tmp_path):
    r, log, inst, _ = run_installer(tmp_path, "minikube/entrypoint.sh", extra_env={"KERNEL_VERSION": "5.10.0-minikube"})
    assert r.returncode == 0, r.stdout + r.stderr
    assert "https://cdn.kernel.org/pub/linux/kernel/v5.x/linux-5.10.tar.xz" in log and "make modules_prepare" in log
    assert f"--kernel-source-path={tmp_path}/ksrc" in log
    assert (tmp_path / "ksrc/include/generated/utsrelease.h").read_text().strip() == '#define UTS_RELEASE "5.10.0-minikube"'


def test_commands_in_manifests_exist_in_the_images_they_run_in():
    """A container command with an absolute path must be something its image's Dockerfile puts there, something the transport
    installer drops into the host mount (/usr/local/nvidia/bin), or a base-image shell."""
    import glob
    import re
    import sys
    sys.path.insert(0, os.path.join(ROOT, "deploy"))
    import generate
    dockerfile_of = {"b200-device-plugin": "device-plugin", "b200-device-plugin-native": "device-plugin-native", "b200-partition-gpu": "partition-gpu",
                     "b200-nri-device-injector": "nri-device-injector", "b200-persistenced": "persistenced", "b200coll-installer": "b200coll-installer",
                     "b200-topology-scheduler": "topology-scheduler", "b200-xid-inject": "xid-inject", "b200-ubuntu-driver-installer": "driver-installer-ubuntu",
                     "b200-minikube-driver-installer": "driver-installer-ubuntu", "fastsocket-installer": "fastsocket-installer"}
    installed_on_host = set(re.findall(r'"\$\{BIN\}/([\w-]+)"|\$\{SRC\}/bin/([\w-]+)" "\$\{BIN\}/"', open(os.path.join(SCRIPTS, "b200coll-install.sh")).read()))
    host_bins = {"/usr/local/nvidia/bin/" + (a or b) for a, b in installed_on_host} | {f"/usr/local/nvidia/bin/{n}_perf" for n in ("all_reduce", "all_gather", "reduce_scatter", "alltoall", "broadcast", "reduce")}
    seen = 0
    for path in glob.glob(os.path.join(ROOT, "deploy", "**", "*.yaml"), recursive=True):
        for doc in yaml.safe_load_all(open(path)):
            stack = [doc]
            while stack:
                o = stack.pop()
                if isinstance(o, list):
                    stack.extend(o)
                elif isinstance(o, dict):
                    stack.extend(o.values())
                    image, cmd = o.get("image"), o.get("command") or []
                    if isinstance(image, str) and image.startswith(generate.REG + "/") and cmd and cmd[0].startswith("/"):
                        name = image.split("/")[-1].split(":")[0]
                        assert name in dockerfile_of, f"{path}: no Dockerfile known for image {name}"
                        dockerfile = open(os.path.join(ROOT, "docker", dockerfile_of[name] + ".Dockerfile")).read()
                        ok = cmd[0] in ("/bin/bash", "/bin/sh", "/bin/true") or cmd[0] in host_bins or re.search(r"^COPY .*\s" + re.escape(cmd[0]) + r"\s*$", dockerfile, re.M) \
                            or re.search(r"^COPY .*\s" + re.escape(os.path.dirname(cmd[0])) + r"/\s*$", dockerfile, re.M)
                        assert ok, f"{os.path.relpath(path, ROOT)}: {cmd[0]} is not provided by docker/{dockerfile_of[name]}.Dockerfile nor by the transport installer"
                        seen += 1
    assert seen >= 8


def _pod_specs(doc):
    kind = doc.get("kind")
    if kind == "Pod":
        yield doc["metadata"], doc["spec"], None
    elif kind in ("DaemonSet", "Deployment", "StatefulSet", "Job"):
        t = doc["spec"]["template"]
        yield t.get("metadata", {}), t["spec"], doc["spec"].get("selector")
    elif kind == "JobSet":
        for rj in doc["spec"]["replicatedJobs"]:
            t = rj["template"]["spec"]["template"]
            yield t.get("metadata", {}), t["spec"], None


def test_structural_validity_of_every_workload():
    """What `kubectl apply --dry-run=server` would catch first: selectors that do not select their own pods, volumeMounts without a
    volume, duplicate container names, probes/ports on unknown names, ConfigMap references nobody defines."""
    import glob
    configmaps, workloads = set(), 0
    docs = []
    for path in sorted(glob.glob(os.path.join(ROOT, "deploy", "**", "*.yaml"), recursive=True)):
        for doc in yaml.safe_load_all(open(path)):
            if isinstance(doc, dict) and "kind" in doc:
                docs.append((os.path.relpath(path, ROOT), doc))
                if doc["kind"] == "ConfigMap":
                    configmaps.add(doc["metadata"]["name"])
    for path, doc in docs:
        assert doc.get("apiVersion") and doc.get("metadata", {}).get("name") or doc["kind"] in ("Kustomization",), path
        for meta, spec, selector in _pod_specs(doc):
            workloads += 1
            if selector:
                for k, v in (selector.get("matchLabels") or {}).items():
                    assert (meta.get("labels") or {}).get(k) == v, f"{path}: selector {k}={v} does not match the pod template labels {meta.get('labels')}"
            volumes = {v["name"] for v in spec.get("volumes") or []}
            # ResourceClaims and generic ephemeral volumes also count as mountable names
            containers = (spec.get("initContainers") or []) + (spec.get("containers") or [])
            names = [c["name"] for c in containers]
            assert len(names) == len(set(names)) and names, f"{path}: container names {names}"
            for c in containers:
                assert c.get("image"), f"{path}: container {c['name']} has no image"
                for m in c.get("volumeMounts") or []:
                    assert m["name"] in volumes, f"{path}: container {c['name']} mounts undeclared volume {m['name']} (have {sorted(volumes)})"
                for e in c.get("env") or []:
                    ref = ((e.get("valueFrom") or {}).get("configMapKeyRef") or {})
                    if ref and not ref.get("optional"):
                        assert ref["name"] in configmaps, f"{path}: env {e['name']} needs ConfigMap {ref['name']} which no manifest defines"
                for e in c.get("envFrom") or []:
                    ref = e.get("configMapRef") or {}
                    if ref and not ref.get("optional"):
                        assert ref["name"] in configmaps, f"{path}: envFrom ConfigMap {ref['name']} undefined"
            for v in spec.get("volumes") or []:
                cm = (v.get("configMap") or {}).get("name")
                if cm and not (v.get("configMap") or {}).get("optional"):
                    assert cm in configmaps, f"{path}: volume {v['name']} needs ConfigMap {cm} which no manifest defines"
            used = {m["name"] for c in containers for m in c.get("volumeMounts") or []}
            assert volumes <= used | {v["name"] for v in spec.get("volumes") or [] if "emptyDir" in v and False}, f"{path}: volumes declared but never mounted: {sorted(volumes - used)}"
    assert workloads >= 50


def test_services_select_a_pod_defined_in_the_same_file_and_critical_priority_only_in_kube_system():
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "deploy", "**", "*.yaml"), recursive=True)):
        docs = [d for d in yaml.safe_load_all(open(path)) if isinstance(d, dict)]
        pods = [(m.get("labels") or {}) for d in docs for m, _, _ in _pod_specs(d)]
        for d in docs:
            if d.get("kind") == "Service" and (d["spec"].get("selector")):
                sel = d["spec"]["selector"]
                assert any(all(lab.get(k) == v for k, v in sel.items()) for lab in pods), f"{os.path.relpath(path, ROOT)}: Service {d['metadata']['name']} selects {sel}, no pod in the file has it"
            for _, spec, _ in _pod_specs(d):
                if spec.get("priorityClassName", "").startswith("system-"):
                    assert d["metadata"].get("namespace") == "kube-system", f"{os.path.relpath(path, ROOT)}: {spec['priorityClassName']} is only admitted in kube-system"


# ------------------------------------------------------------------------------------------------- remaining host scripts
def test_cos_enable_kdump_is_idempotent_and_cos_only(tmp_path):
    osr = tmp_path / "os-release"
    script = os.path.join(SCRIPTS, "cos-enable-kdump.sh")
    osr.write_text("ID=ubuntu\n")
    env = {**os.environ, "OS_RELEASE": str(osr), "KDUMP_HELPER": stub(tmp_path, "kdump_helper"), "REBOOT": stub(tmp_path, "reboot")}
    assert subprocess.run(["bash", script], env=env).returncode == 0 and not (tmp_path / "calls.log").exists()          # not COS: untouched
    osr.write_text("ID=cos\n")
    env["KDUMP_HELPER"] = stub(tmp_path, "kdump_helper", 'if [ "$1" = status ]; then echo "kdump is ready"; fi')
    assert subprocess.run(["bash", script], env=env).returncode == 0
    assert "reboot" not in (tmp_path / "calls.log").read_text()                                                          # already enabled: no reboot
    (tmp_path / "calls.log").unlink()
    env["KDUMP_HELPER"] = stub(tmp_path, "kdump_helper", 'if [ "$1" = status ]; then echo "kdump is not enabled"; fi')
    assert subprocess.run(["bash", script], env=env).returncode == 0
    log = (tmp_path / "calls.log").read_text()
    assert "kdump_helper enable" in log and log.strip().endswith("reboot")                                              # enable, then one reboot


def test_fix_hostname_writes_the_networkd_profile(tmp_path):
    net = tmp_path / "network"; net.mkdir()
    env = {**os.environ, "NETWORKD_DIR": str(net), "CURL": stub(tmp_path, "curl", "echo gke-a4-node-7.c.project.internal"), "HOSTNAMECTL": stub(tmp_path, "hostnamectl"),
           "NETWORKCTL": stub(tmp_path, "networkctl")}
    assert subprocess.run(["bash", os.path.join(SCRIPTS, "fix-hostname.sh")], env=env).returncode == 0
    prof = (net / "97-temp.network").read_text()
    assert "Name=!eth0" in prof and "UseHostname=false" in prof and "DHCP=yes" in prof
    log = (tmp_path / "calls.log").read_text()
    assert "hostnamectl set-hostname gke-a4-node-7" in log and "networkctl reload" in log and "Metadata-Flavor: Google" in log


def test_vgpu_machine_type_file(tmp_path):
    env = {**os.environ, "ROOT_MOUNT_DIR": str(tmp_path / "root"), "CURL": stub(tmp_path, "curl", "echo projects/123/machineTypes/g4-standard-12")}
    assert subprocess.run(["bash", os.path.join(SCRIPTS, "vgpu-machine-type.sh")], env=env).returncode == 0
    assert (tmp_path / "root/etc/nvidia/machine_type.txt").read_text().strip() == "g4-standard-12"      # what b200-persistenced's gridd decision reads


def test_asapd_lite_supervisor_restarts_on_dad_failure(tmp_path):
    daemon_pid = tmp_path / "daemon.pid"
    run = stub(tmp_path, "run_asapd_lite.sh", f"echo $$ > {daemon_pid}; exec sleep 30")
    ip = stub(tmp_path, "ip", 'case "$*" in "-o link show") echo "1: lo: <UP>"; echo "2: gpu0rdma0: <UP>";; "-6 addr show dev gpu0rdma0") echo "inet6 fe80::1/64 scope link dadfailed tentative";; esac')
    env = {**os.environ, "ASAPD_RUN": run, "IP": ip, "ASAPD_POLL_S": "0.1"}
    r = subprocess.run(["bash", os.path.join(SCRIPTS, "asapd-lite-run.sh")], env=env, capture_output=True, text=True, timeout=30)
    assert r.returncode == 1 and "DAD failed on gpu0rdma0" in r.stdout                                   # exit 1 => kubelet restarts the pod
    log = (tmp_path / "calls.log").read_text()
    assert "ip link set dev gpu0rdma0 down" in log and "ip link set dev gpu0rdma0 up" in log and "--enable-hairpin-probe" in log
    ip = stub(tmp_path, "ip", 'case "$*" in "-o link show") echo "2: gpu0rdma0: <UP>";; *) echo "inet6 fe80::1/64 scope link";; esac')
    run = stub(tmp_path, "run_asapd_lite.sh", "sleep 0.3; exit 7")
    r = subprocess.run(["bash", os.path.join(SCRIPTS, "asapd-lite-run.sh")], env={**env, "ASAPD_RUN": run, "IP": ip}, capture_output=True, text=True, timeout=30)
    assert r.returncode == 7                                                                             # healthy links: the daemon's own exit status is propagated


def test_run_nccl_builds_the_reference_mpirun_line(tmp_path):
    """run-nccl.sh <bench> <ld path> <gpus/node> <nics> <min> <max> <nhosts>: np = gpus x hosts, TCPX env, nccl-tests flags (S12)."""
    fake_bin = tmp_path / "bin"; fake_bin.mkdir()
    for tool in ("mpirun", "taskset"):
        (fake_bin / tool).write_text(f"#!/bin/bash\necho \"{tool} $*\" >> {tmp_path}/calls.log\n"); (fake_bin / tool).chmod(0o755)
    env = {**os.environ, "PATH": f"{fake_bin}:{os.environ['PATH']}"}
    r = subprocess.run(["bash", os.path.join(SCRIPTS, "run-nccl.sh"), "all_gather_perf", "/usr/local/nvidia/lib64", "8", "eth1,eth2,eth3,eth4", "1M", "512M", "2"], env=env,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    line = (tmp_path / "calls.log").read_text()
    assert "-np 16 " in line and "--hostfile /scripts/hostfiles2/hostfile8" in line and "NCCL_GPUDIRECTTCPX_SOCKET_IFNAME=eth1,eth2,eth3,eth4" in line
    assert "NCCL_ALGO=Ring" in line and "NCCL_PROTO=Simple" in line and "NCCL_BUFFSIZE=4194304" in line
    assert line.rstrip().endswith("all_gather_perf -b 1M -e 512M -f 2 -g 1 -w 5 --iters 100 -c 0")


def test_jobset_worker_head_and_followers(tmp_path):
    """Head: waits for every peer, writes `host slots=N` lines, runs mpirun with np = nodes x slots and the NCCL_* environment.
    Follower: stays while the head answers ssh, exits 0 once it is gone (reference nccl-test-a4x-max-jobset.yaml:104-163)."""
    script = os.path.join(SCRIPTS, "jobset-worker.sh")
    base = {**os.environ, "NUM_NODES": "2", "GPUS_PER_NODE": "4", "JOBSET_NAME": "nccl", "SSHD_START": "true", "POLL_S": "0.05", "HOSTFILE": str(tmp_path / "hostfile"),
            "NCCL_ENV_SCRIPT": str(tmp_path / "absent.sh"), "NCCL_DEBUG": "INFO"}
    # head: the second peer answers only from the third attempt on
    tries = tmp_path / "tries"
    ssh = stub(tmp_path, "ssh", f'case "$*" in *nccl-w-0-1*) n=$(cat {tries} 2>/dev/null || echo 0); echo $((n+1)) > {tries}; [ "$n" -ge 2 ];; esac')
    r = subprocess.run(["bash", script], env={**base, "JOB_COMPLETION_INDEX": "0", "SSH": ssh, "MPIRUN": stub(tmp_path, "mpirun")}, capture_output=True, text=True, timeout=30)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "hostfile").read_text().splitlines() == ["nccl-w-0-0.nccl slots=4", "nccl-w-0-1.nccl slots=4"]
    assert "waiting for nccl-w-0-1.nccl" in r.stdout
    line = [l for l in (tmp_path / "calls.log").read_text().splitlines() if l.startswith("mpirun")][0]
    assert "-np 8 " in line and "-x NCCL_DEBUG" in line and "-x NCCL_TESTS_SPLIT_MASK=0x0" in line and line.endswith("all_gather_perf -b 1K -e 8G -f 2 -g 1 -w 5 --iters 100 -c 1")
    # follower: head reachable for the first 4 probes, then gone
    (tmp_path / "calls.log").unlink(); tries.write_text("0")
    ssh = stub(tmp_path, "ssh", f'n=$(cat {tries}); echo $((n+1)) > {tries}; [ "$n" -lt 4 ]')
    r = subprocess.run(["bash", script], env={**base, "JOB_COMPLETION_INDEX": "1", "SSH": ssh, "MPIRUN": stub(tmp_path, "mpirun")}, capture_output=True, text=True, timeout=30)
    assert r.returncode == 0
    log = (tmp_path / "calls.log").read_text()
    assert log.count("ssh ") == 5 and "mpirun" not in log           # 4 successful probes + the failing one; followers never launch anything
