#!/usr/bin/env python
"""Single source of truth for every Kubernetes manifest under deploy/.

The reference keeps ~74 hand-edited YAML files whose variants drift apart. Here each family is described
once and its variants are derived: `python deploy/generate.py` rewrites deploy/**.yaml, `--check` fails if the tree is stale
(tests/test_manifests.py runs it). Inline shell programs live as real, lint-able files in
deploy/scripts/ and are embedded verbatim.

Families (reference directories in brackets):
  device-plugin/        [cmd/nvidia_gpu, demo/, test/nvidia_gpu]   plugin DaemonSet, RBAC, xid-config, health demo, e2e fixture
  partition-gpu/        [partition_gpu]                            MIG partitioner
  nri-device-injector/  [nri_device_injector]                      standard + Autopilot
  driver-installer/     [nvidia-driver-installer, daemonset.yaml]  cos x7 (+kustomization), ubuntu x6, minikube, legacy root
  transport/            [fast-socket-installer, gpudirect-*, asapd-lite-installer]  b200coll (new) + compat installers + host tweaks
  nccl-test/            [gpudirect-*/nccl-test*.yaml, nccl-config] b200coll single-box pod + per-transport pod pairs + JobSet
  topology-scheduler/   [gke-topology-scheduler, gpudirect-tcpxo/topology-scheduler]
  demo/, example/       [demo, example]                            serving + HPA, training sweep, TPU, minikube, prepull, notebook, gpu-error, MPS
  test/                 [test/nvidia_gpu]                          e2e fixtures
"""
from __future__ import annotations

import argparse
import copy
import json
import sys
from pathlib import Path

import yaml

HERE = Path(__file__).resolve().parent
SCRIPTS = HERE / "scripts"

# ---------------------------------------------------------------------------------------------------- images
REG = "ghcr.io/b200-node-accelerators"
VERSION = (HERE.parent / "VERSION").read_text().strip() if (HERE.parent / "VERSION").exists() else "v0.1.0"   # one version for images, manifests and the Makefile
IMG = {
    "device-plugin": f"{REG}/b200-device-plugin:{VERSION}",
    "device-plugin-native": f"{REG}/b200-device-plugin-native:{VERSION}",
    "partition-gpu": f"{REG}/b200-partition-gpu:{VERSION}",
    "nri-injector": f"{REG}/b200-nri-device-injector:{VERSION}",
    "persistenced": f"{REG}/b200-persistenced:{VERSION}",
    "b200coll": f"{REG}/b200coll-installer:{VERSION}",
    "scheduler": f"{REG}/b200-topology-scheduler:{VERSION}",
    "xid-inject": f"{REG}/b200-xid-inject:{VERSION}",
    "ubuntu-installer": f"{REG}/b200-ubuntu-driver-installer:{VERSION}",
    "minikube-installer": f"{REG}/b200-minikube-driver-installer:{VERSION}",
    "fastsocket": f"{REG}/fastsocket-installer:{VERSION}",
    # third-party payloads pulled exactly as upstream ships them
    "pause": "gke.gcr.io/pause:3.8@sha256:880e63f94b145e46f1b1082bb71b85e21f16b99b180b9996407d61240ceb9830",
    "cos-installer": "cos-nvidia-installer:fixed",
    "gke-ubuntu-installer": "gke-nvidia-installer:fixed",
    "cos-gpu-installer-legacy": "gcr.io/cos-cloud/cos-gpu-installer@sha256:af09af53cb7c6ddce9c96968d5367253684d08108a414c8a1d70f96188427949",
    "tcpx-plugin": "us-docker.pkg.dev/gce-ai-infra/gpudirect-tcpx/nccl-plugin-gpudirecttcpx-dev:v3.1.9",
    "tcpx-rxdm": "us-docker.pkg.dev/gce-ai-infra/gpudirect-tcpx/tcpgpudmarxd-dev:v2.0.12",
    "tcpx-rxdm-old": "us-docker.pkg.dev/gce-ai-infra/gpudirect-tcpx/tcpgpudmarxd-dev:v2.0.9",
    "tcpx-metrics": "us-docker.pkg.dev/gce-ai-infra/gpudirect-tcpx/tcpx-metrics:latest",
    "tcpxo-plugin": "us-docker.pkg.dev/gce-ai-infra/gpudirect-tcpxo/nccl-plugin-gpudirecttcpx-dev:v1.0.15",
    "tcpxo-rxdm": "us-docker.pkg.dev/gce-ai-infra/gpudirect-tcpxo/tcpgpudmarxd-dev:v1.0.21",
    "tcpxo-rxdm-old": "us-docker.pkg.dev/gce-ai-infra/gpudirect-tcpxo/tcpgpudmarxd-dev:v1.0.17",
    "gib": "us-docker.pkg.dev/gce-ai-infra/gpudirect-gib/nccl-plugin-gib:v1.1.1",
    "gib-arm64": "us-docker.pkg.dev/gce-ai-infra/gpudirect-gib/nccl-plugin-gib-arm64:v1.1.0",
    "gib-diag": "us-docker.pkg.dev/gce-ai-infra/gpudirect-gib/nccl-plugin-gib-diagnostic:v1.1.1",
    "gib-diag-arm64": "us-docker.pkg.dev/gce-ai-infra/gpudirect-gib/nccl-plugin-gib-diagnostic-arm64:v1.1.0",
    "gib-a4x-max": "us-docker.pkg.dev/gce-ai-infra/gpudirect-gib/nccl-gib-a4x-max-arm64:v1.1.1",
    "asapd-lite": "us-docker.pkg.dev/gce-ai-infra/asapd-lite/asapd-lite:v0.0.7",
    "distroless-bash": "gke.gcr.io/gke-distroless/bash",
}

NVIDIA_HOST = "/home/kubernetes/bin/nvidia"
NVIDIA_CTR = "/usr/local/nvidia"
ACCEL_KEY = "cloud.google.com/gke-accelerator"


def script(name: str) -> str:
    return (SCRIPTS / name).read_text()


# ---------------------------------------------------------------------------------------------------- building blocks
def host_vol(name: str, path: str, type_: str | None = None) -> dict:
    hp = {"path": path}
    if type_:
        hp["type"] = type_
    return {"name": name, "hostPath": hp}


def mount(name: str, path: str, **kw) -> dict:
    return {"name": name, "mountPath": path, **kw}


def env(**kv) -> list:
    return [{"name": k, "value": str(v)} for k, v in kv.items()]


def node_affinity(*exprs) -> dict:
    return {"nodeAffinity": {"requiredDuringSchedulingIgnoredDuringExecution": {"nodeSelectorTerms": [{"matchExpressions": list(exprs)}]}}}


def expr(key: str, op: str, values: list | None = None) -> dict:
    e = {"key": key, "operator": op}
    if values:
        e["values"] = values
    return e


def pause_container() -> dict:
    return {"name": "pause", "image": IMG["pause"]}


def daemonset(name: str, *, namespace: str = "kube-system", init: list | None = None, containers: list | None = None, volumes: list | None = None, affinity: dict | None = None,
              host_network: bool = True, host_pid: bool = True, critical: bool = True, tolerate_all: bool = True, labels: dict | None = None, extra_spec: dict | None = None,
              annotations: dict | None = None) -> dict:
    """The common node-agent shape: a privileged initContainer does the work, `pause` keeps the pod
    Running, tolerate everything, system-node-critical, RollingUpdate."""
    lab = {"k8s-app": name, **(labels or {})}
    pod_spec: dict = {}
    if critical:
        pod_spec["priorityClassName"] = "system-node-critical"
    if affinity:
        pod_spec["affinity"] = affinity
    if tolerate_all:
        pod_spec["tolerations"] = [{"operator": "Exists"}]
    if host_network:
        pod_spec["hostNetwork"] = True
    if host_pid:
        pod_spec["hostPID"] = True
    if volumes:          # variants share volume lists: keep only what some container of this variant mounts
        mounted = {m["name"] for c in (init or []) + (containers or []) for m in c.get("volumeMounts") or []}
        volumes = [v for v in volumes if v["name"] in mounted]
    if volumes:
        pod_spec["volumes"] = volumes
    if init:
        pod_spec["initContainers"] = init
    pod_spec["containers"] = containers or [pause_container()]
    pod_spec.update(extra_spec or {})
    meta = {"labels": {"name": name, **lab}}
    if annotations:
        meta["annotations"] = annotations
    return {"apiVersion": "apps/v1", "kind": "DaemonSet", "metadata": {"name": name, "namespace": namespace, "labels": lab},
            "spec": {"selector": {"matchLabels": {"k8s-app": name}}, "updateStrategy": {"type": "RollingUpdate"}, "template": {"metadata": meta, "spec": pod_spec}}}


def bash(script_text: str) -> dict:
    return {"command": ["bash", "-c", script_text]}


# ---------------------------------------------------------------------------------------------------- device plugin
def device_plugin_ds(*, name="b200-gpu-device-plugin", health=True, metrics=True, mps=False, selector_key=ACCEL_KEY, transport_note=True) -> dict:
    args = ["python", "-m", "container_engine_accelerators_b200.agent.main"]
    if metrics:
        args.append("--enable-container-gpu-metrics")
    if health:
        args.append("--enable-health-monitoring")
    args.append("--publish-driver-version")
    vols = [host_vol("device-plugin", "/var/lib/kubelet/device-plugins", "Directory"), host_vol("dev", "/dev", "Directory"), host_vol("nvidia", NVIDIA_HOST, "Directory"),
            host_vol("pod-resources", "/var/lib/kubelet/pod-resources", "Directory"), host_vol("proc", "/proc", "Directory"), host_vol("nvidia-config", "/etc/nvidia", "DirectoryOrCreate"),
            {"name": "shm", "hostPath": {"path": "/dev/shm", "type": "Directory"}}]
    mounts = [mount("device-plugin", "/device-plugin"), mount("dev", "/dev"), mount("nvidia", NVIDIA_CTR), mount("pod-resources", "/var/lib/kubelet/pod-resources"), mount("proc", "/proc"),
              mount("nvidia-config", "/etc/nvidia"), mount("shm", "/dev/shm")]
    if mps:
        vols.append(host_vol("mps", "/tmp/nvidia-mps", "DirectoryOrCreate"))
        mounts.append(mount("mps", "/tmp/nvidia-mps"))
    ctr = {"name": "b200-gpu-device-plugin", "image": IMG["device-plugin"], "command": args,
           "env": [{"name": "XID_CONFIG", "valueFrom": {"configMapKeyRef": {"name": "xid-config", "key": "HealthCriticalXid", "optional": True}}},
                   {"name": "NODE_NAME", "valueFrom": {"fieldRef": {"fieldPath": "spec.nodeName"}}}] + env(LD_LIBRARY_PATH=f"{NVIDIA_CTR}/lib64"),
           "ports": [{"containerPort": 2112, "name": "metrics"}], "resources": {"requests": {"cpu": "50m", "memory": "100Mi"}, "limits": {"memory": "200Mi"}},
           "securityContext": {"privileged": True}, "volumeMounts": mounts}
    ds = daemonset(name, containers=[ctr], volumes=vols, affinity=node_affinity(expr(selector_key, "Exists")), host_network=False, host_pid=False, tolerate_all=False,
                   extra_spec={"serviceAccountName": "gpu-device-plugin", "restartPolicy": "Always", "securityContext": {"seccompProfile": {"type": "RuntimeDefault"}},
                               "tolerations": [{"effect": "NoExecute", "operator": "Exists"}, {"effect": "NoSchedule", "operator": "Exists"}]})
    return ds


def device_plugin_native_ds() -> dict:
    """Native variant: one C++ binary owns the kubelet socket, metrics, Unhealthy marking and the Kubernetes-side status
    (Xid Events, the XidCriticalError Node condition, driver-version annotations; agent/native/dp/kube.hpp). No Python in the pod."""
    ds = device_plugin_ds(name="b200-gpu-device-plugin-native")
    spec = ds["spec"]["template"]["spec"]
    native = spec["containers"][0]
    native.update({"name": "b200-device-plugin", "image": IMG["device-plugin-native"],
                   "command": ["/usr/bin/b200-device-plugin", "-logtostderr", "-enable-container-gpu-metrics", "-enable-health-monitoring", "-publish-driver-version"],
                   "resources": {"requests": {"cpu": "50m", "memory": "20Mi"}, "limits": {"memory": "100Mi"}}})
    spec["containers"] = [native]
    return ds


def device_plugin_rbac() -> list:
    lab = {"k8s-app": "gpu-device-plugin"}
    return [
        {"apiVersion": "v1", "kind": "ServiceAccount", "metadata": {"name": "gpu-device-plugin", "namespace": "kube-system", "labels": lab}},
        {"apiVersion": "rbac.authorization.k8s.io/v1", "kind": "ClusterRole", "metadata": {"name": "gpu-device-plugin", "labels": {**lab, "addonmanager.kubernetes.io/mode": "Reconcile"}},
         "rules": [{"apiGroups": [""], "resources": ["nodes"], "verbs": ["update", "patch", "get", "list", "watch"]},
                   {"apiGroups": [""], "resources": ["nodes/status"], "verbs": ["update", "patch", "get", "list", "watch"]},
                   {"apiGroups": [""], "resources": ["events"], "verbs": ["create", "patch", "update"]}]},
        {"apiVersion": "rbac.authorization.k8s.io/v1", "kind": "ClusterRoleBinding", "metadata": {"name": "gpu-device-plugin", "labels": lab},
         "roleRef": {"apiGroup": "rbac.authorization.k8s.io", "kind": "ClusterRole", "name": "gpu-device-plugin"},
         "subjects": [{"kind": "ServiceAccount", "name": "gpu-device-plugin", "namespace": "kube-system"}]},
    ]


def xid_config(values: str = "31,48,79") -> dict:
    return {"apiVersion": "v1", "kind": "ConfigMap", "metadata": {"name": "xid-config", "namespace": "kube-system"}, "data": {"HealthCriticalXid": values}}


def gpu_config_map(partition: str = "", transport: str = "b200coll", sharing: dict | None = None) -> dict:
    cfg: dict = {}
    if partition:
        cfg["GPUPartitionSize"] = partition
    if sharing:
        cfg["GPUSharingConfig"] = sharing
    if transport:
        cfg["Transport"] = {"Name": transport}
    return {"apiVersion": "v1", "kind": "ConfigMap", "metadata": {"name": "gpu-config", "namespace": "kube-system"}, "data": {"gpu_config.json": json.dumps(cfg, indent=1)}}


# ---------------------------------------------------------------------------------------------------- partitioner / persistenced / NRI
def partition_container() -> dict:
    return {"name": "partition-gpus", "image": IMG["partition-gpu"], "command": ["/usr/bin/b200-partition-gpu", "-logtostderr"], "env": env(LD_LIBRARY_PATH=f"{NVIDIA_CTR}/lib64"),
            "resources": {"requests": {"cpu": "150m"}}, "securityContext": {"privileged": True},
            "volumeMounts": [mount("nvidia-install-dir-host", NVIDIA_CTR), mount("dev", "/dev"), mount("nvidia-config", "/etc/nvidia")]}


def partition_gpu_ds() -> dict:
    vols = [host_vol("dev", "/dev"), host_vol("nvidia-install-dir-host", NVIDIA_HOST), host_vol("nvidia-config", "/etc/nvidia")]
    return daemonset("b200-partition-gpus", init=[partition_container()], volumes=vols, affinity=node_affinity(expr(ACCEL_KEY, "Exists")))


def persistenced_container(restart_always: bool = False) -> dict:
    c = {"name": "nvidia-daemon-installer", "image": IMG["persistenced"], "command": ["/usr/bin/b200-persistenced", "-logtostderr"], "env": env(ROOT_MOUNT_DIR="/root"),
         "resources": {"requests": {"cpu": "50m"}}, "securityContext": {"privileged": True},
         "volumeMounts": [mount("nvidia-install-dir-host", NVIDIA_CTR), mount("dev", "/dev"), mount("nvidia-config", "/etc/nvidia"), mount("root-mount", "/root")]}
    if restart_always:
        c["restartPolicy"] = "Always"
    return c


def nri_injector_ds(autopilot: bool = False) -> dict:
    ns = "gpudirect-system" if autopilot else "kube-system"
    ctr = {"name": "device-injector", "image": IMG["nri-injector"], "command": ["/usr/bin/b200-nri-device-injector", "--idx", "10"],   # native binary; `python -m ...agent.nri` is the equivalent
           "resources": {"requests": {"cpu": "10m", "memory": "50Mi"}, **({"limits": {"cpu": "100m", "memory": "100Mi"}} if autopilot else {})},
           "securityContext": {"privileged": True}, "volumeMounts": [mount("nri-socket", "/var/run/nri"), mount("dev", "/dev")]}
    aff = node_affinity(expr(ACCEL_KEY, "In", ["nvidia-b200", "nvidia-h100-80gb", "nvidia-h100-mega-80gb", "nvidia-rtx-pro-6000"]))
    return daemonset("b200-nri-device-injector", namespace=ns, containers=[ctr], volumes=[host_vol("nri-socket", "/var/run/nri"), host_vol("dev", "/dev")], affinity=aff,
                     host_network=False, host_pid=False, critical=not autopilot)


# ---------------------------------------------------------------------------------------------------- driver installers
def cos_installer_env() -> list:
    return env(NVIDIA_INSTALL_DIR_HOST=NVIDIA_HOST, NVIDIA_INSTALL_DIR_CONTAINER=NVIDIA_CTR, VULKAN_ICD_DIR_HOST=f"{NVIDIA_HOST}/vulkan/icd.d", VULKAN_ICD_DIR_CONTAINER="/etc/vulkan/icd.d",
               ROOT_MOUNT_DIR="/root", COS_TOOLS_DIR_HOST="/var/lib/cos-tools", COS_TOOLS_DIR_CONTAINER="/build/cos-tools")


COS_VOLS = [host_vol("dev", "/dev"), host_vol("vulkan-icd-mount", f"{NVIDIA_HOST}/vulkan/icd.d"), host_vol("nvidia-install-dir-host", NVIDIA_HOST), host_vol("root-mount", "/"),
            host_vol("cos-tools", "/var/lib/cos-tools"), host_vol("nvidia-config", "/etc/nvidia")]
COS_MOUNTS = [mount("nvidia-install-dir-host", NVIDIA_CTR), mount("vulkan-icd-mount", "/etc/vulkan/icd.d"), mount("dev", "/dev"), mount("root-mount", "/root"), mount("cos-tools", "/build/cos-tools")]


def cos_driver_ds(variant: str) -> dict:
    """variant: preloaded | preloaded-latest | preloaded-latest-a4x | confidential | confidential-latest | nvidia-mig | vgpu-latest"""
    latest = variant.endswith("latest") or variant.endswith("a4x")
    version_flag = "--version=latest" if latest else ""
    installer = {"name": "nvidia-driver-installer", "image": IMG["cos-installer"], "imagePullPolicy": "Never", "resources": {"requests": {"cpu": "150m"}}, "securityContext": {"privileged": True},
                 "env": cos_installer_env() + env(COS_GPU_INSTALLER_VERSION_FLAG=version_flag), "volumeMounts": list(COS_MOUNTS)}
    exprs = [expr(ACCEL_KEY, "Exists"), expr("cloud.google.com/gke-gpu-driver-version", "DoesNotExist")]
    init, containers = [], [pause_container()]
    if variant.startswith("confidential"):
        installer.update(bash(script("cos-confidential-install.sh")))
        exprs.append(expr("cloud.google.com/gke-confidential-nodes-instance-type", "In", ["TDX", "SEV"]))
        init = [installer, persistenced_container(restart_always=True), partition_container()]
    elif variant == "vgpu-latest":
        mt = {"name": "machine-type", "image": IMG["distroless-bash"], "securityContext": {"privileged": True}, "volumeMounts": [mount("root-mount", "/root")], **bash(script("vgpu-machine-type.sh"))}
        installer.update(bash(script("cos-driver-install.sh")))
        exprs.append(expr("cloud.google.com/gke-confidential-nodes-instance-type", "DoesNotExist"))
        init = [mt, installer]
        containers = [persistenced_container(), pause_container()]     # gridd launcher runs as a regular container here instead of the MIG step
    else:
        installer.update(bash(script("cos-driver-install.sh")))
        exprs.append(expr("cloud.google.com/gke-confidential-nodes-instance-type", "DoesNotExist"))
        init = [installer]
        if variant == "preloaded-latest-a4x":
            exprs.append(expr("node.kubernetes.io/instance-type", "In", ["a4x-highgpu-4g", "a4x-highgpu-4g-nolssd"]))    # GB200: no MIG step
        else:
            init.append(partition_container())
    name = "nvidia-driver-installer" if variant != "nvidia-mig" else "nvidia-driver-installer-mig"
    return daemonset(name, init=init, containers=containers, volumes=copy.deepcopy(COS_VOLS), affinity=node_affinity(*exprs))


def ubuntu_driver_ds(pin: str = "", preloaded: bool = True) -> dict:
    image = IMG["gke-ubuntu-installer"] if preloaded else IMG["ubuntu-installer"]
    e = env(NVIDIA_INSTALL_DIR_HOST=NVIDIA_HOST, NVIDIA_INSTALL_DIR_CONTAINER=NVIDIA_CTR, VULKAN_ICD_DIR_HOST=f"{NVIDIA_HOST}/vulkan/icd.d", VULKAN_ICD_DIR_CONTAINER="/etc/vulkan/icd.d", ROOT_MOUNT_DIR="/root")
    if pin:
        e += env(NVIDIA_DRIVER_VERSION=pin)
    c = {"name": "nvidia-driver-installer", "image": image, "resources": {"requests": {"cpu": "150m"}}, "securityContext": {"privileged": True}, "env": e,
         "volumeMounts": [mount("nvidia-install-dir-host", NVIDIA_CTR), mount("vulkan-icd-mount", "/etc/vulkan/icd.d"), mount("dev", "/dev"), mount("root-mount", "/root")]}
    if preloaded:
        c["imagePullPolicy"] = "Never"
    vols = [host_vol("dev", "/dev"), host_vol("vulkan-icd-mount", f"{NVIDIA_HOST}/vulkan/icd.d"), host_vol("nvidia-install-dir-host", NVIDIA_HOST), host_vol("root-mount", "/"), host_vol("nvidia-config", "/etc/nvidia")]
    return daemonset("nvidia-driver-installer-ubuntu", init=[c, partition_container()], volumes=vols,
                     affinity=node_affinity(expr(ACCEL_KEY, "Exists"), expr("cloud.google.com/gke-gpu-driver-version", "DoesNotExist")))


def minikube_driver_ds() -> dict:
    c = {"name": "nvidia-driver-installer", "image": IMG["minikube-installer"], "resources": {"requests": {"cpu": "150m"}}, "securityContext": {"privileged": True},
         "env": env(NVIDIA_INSTALL_DIR_HOST=NVIDIA_HOST, NVIDIA_INSTALL_DIR_CONTAINER=NVIDIA_CTR, ROOT_MOUNT_DIR="/root"),
         "volumeMounts": [mount("nvidia-install-dir-host", NVIDIA_CTR), mount("dev", "/dev"), mount("root-mount", "/root")]}
    return daemonset("nvidia-driver-installer-minikube", init=[c], volumes=[host_vol("dev", "/dev"), host_vol("nvidia-install-dir-host", NVIDIA_HOST), host_vol("root-mount", "/")], affinity=None)


def legacy_root_ds() -> dict:
    c = {"name": "nvidia-driver-installer", "image": IMG["cos-gpu-installer-legacy"], "resources": {"requests": {"cpu": "150m"}}, "securityContext": {"privileged": True},
         "env": cos_installer_env(), "volumeMounts": [mount("nvidia-install-dir-host", NVIDIA_CTR), mount("dev", "/dev"), mount("root-mount", "/root")]}
    return daemonset("nvidia-driver-installer-legacy", init=[c], volumes=[host_vol("dev", "/dev"), host_vol("nvidia-install-dir-host", NVIDIA_HOST), host_vol("root-mount", "/")],
                     affinity=node_affinity(expr(ACCEL_KEY, "Exists")))


# ---------------------------------------------------------------------------------------------------- transport installers
def b200coll_installer_ds(autopilot: bool = False) -> dict:
    ns = "b200coll-system" if autopilot else "kube-system"
    c = {"name": "b200coll-installer", "image": IMG["b200coll"], "resources": {"requests": {"cpu": "150m"}, **({"limits": {"cpu": "500m", "memory": "256Mi"}} if autopilot else {})},
         "securityContext": {"privileged": True}, "env": env(NCCL_INSTALL_DIR=f"{NVIDIA_CTR}/lib64", B200COLL_BIN_DIR=f"{NVIDIA_CTR}/bin", LD_LIBRARY_PATH=f"{NVIDIA_CTR}/lib64"),
         "volumeMounts": [mount("library-dir-host", NVIDIA_CTR), mount("dev", "/dev")], **bash(script("b200coll-install.sh"))}
    return daemonset("b200coll-installer", namespace=ns, init=[c], volumes=[host_vol("library-dir-host", NVIDIA_HOST), host_vol("dev", "/dev")],
                     affinity=node_affinity(expr(ACCEL_KEY, "In", ["nvidia-b200"])), critical=not autopilot)


def compat_installer_ds(kind: str, autopilot: bool = False, arm64: bool = False) -> dict:
    """fast-socket | tcpx | tcpxo | rdma — third-party NCCL net plugins for the inter-node path (unchanged payloads)."""
    ns = {"tcpx": "gpudirect-system", "tcpxo": "gpudirect-system", "rdma": "rdma-system"}.get(kind, "kube-system") if autopilot else "kube-system"
    vols = [host_vol("library-dir-host", NVIDIA_HOST)]
    mounts = [mount("library-dir-host", NVIDIA_CTR)]
    init = []
    if kind == "fast-socket":
        c = {"name": "fast-socket-installer", "image": IMG["fastsocket"], "env": env(NCCL_INSTALL_DIR=f"{NVIDIA_CTR}/lib64"), "volumeMounts": mounts,
             **bash("cp /usr/lib/libnccl-net.so ${NCCL_INSTALL_DIR}/\n")}
        aff = node_affinity(expr("cloud.google.com/gke-nccl-fastsocket", "Exists"))
        init = [c]
    else:
        image = {"tcpx": IMG["tcpx-plugin"], "tcpxo": IMG["tcpxo-plugin"], "rdma": IMG["gib-arm64"] if arm64 else IMG["gib"]}[kind]
        src = {"tcpx": "/var/lib/tcpx/lib64", "tcpxo": "/var/lib/tcpxo/lib64", "rdma": "/usr/local/gib/lib64"}[kind]
        e = env(TRANSPORT_SRC_DIR=src, NCCL_INSTALL_DIR=f"{NVIDIA_CTR}/lib64")
        if kind in ("tcpx", "tcpxo"):
            vols.append(host_vol(kind, f"/var/lib/{kind}"))
            mounts = mounts + [mount(kind, f"/var/lib/{kind}")]
        if kind == "rdma":
            vols.append(host_vol("gib", "/home/kubernetes/bin/gib"))
            mounts = mounts + [mount("gib", "/usr/local/gib-host")]
            e += env(TRANSPORT_EXTRA_SRC="/usr/local/gib", TRANSPORT_EXTRA_DST="/usr/local/gib-host")
        c = {"name": f"nccl-{kind}-installer", "image": image, "resources": {"requests": {"cpu": "150m"}}, "securityContext": {"privileged": True}, "env": e, "volumeMounts": mounts,
             **bash(script("transport-install.sh"))}
        if kind == "tcpxo":
            prep = {"name": "tcpxo-host-prep", "image": IMG["tcpxo-plugin"], "securityContext": {"privileged": True},
                    "command": ["nsenter", "-at", "1", "--", "bash", "-c", script("tcpxo-host-prep.sh")]}
            vols.append(host_vol("aperture-devices", "/dev/aperture_devices"))
            init = [prep, c]
        else:
            init = [c]
        accel = {"tcpx": ["nvidia-h100-80gb"], "tcpxo": ["nvidia-h100-mega-80gb"], "rdma": ["nvidia-gb200"] if arm64 else ["nvidia-h200-141gb", "nvidia-b200"]}[kind]
        exprs = [expr(ACCEL_KEY, "In", accel)]
        if arm64:
            exprs.append(expr("kubernetes.io/arch", "In", ["arm64"]))
        aff = node_affinity(*exprs)
    name = f"nccl-{kind}-installer" + ("-a4x" if arm64 else "")
    return daemonset(name, namespace=ns, init=init, volumes=vols, affinity=aff, critical=not autopilot)


def host_tweak_ds(name: str, script_name: str, accel: list | None = None, extra_affinity: list | None = None, nsenter: bool = True) -> dict:
    cmd = (["nsenter", "-at", "1", "--", "bash", "-c", script(script_name)] if nsenter else ["bash", "-c", script(script_name)])
    c = {"name": name, "image": IMG["distroless-bash"], "securityContext": {"privileged": True}, "command": cmd}
    exprs = [expr(ACCEL_KEY, "In", accel)] if accel else [expr("cloud.google.com/gke-gpu", "In", ["true"])]
    return daemonset(name, init=[c], affinity=node_affinity(*(exprs + (extra_affinity or []))))


def asapd_lite_ds() -> dict:
    c = {"name": "asapd-lite", "image": IMG["asapd-lite"], "securityContext": {"privileged": True}, "resources": {"limits": {"hugepages-2Mi": "8Gi", "memory": "8Gi"}},
         "livenessProbe": {"httpGet": {"path": "/healthz", "port": 19540}, "initialDelaySeconds": 30, "periodSeconds": 10},
         "readinessProbe": {"httpGet": {"path": "/healthz", "port": 19540}, "periodSeconds": 10}, **bash(script("asapd-lite-run.sh"))}
    return daemonset("asapd-lite-installer", containers=[c], affinity=node_affinity(expr("node.kubernetes.io/instance-type", "In", ["a4x-maxgpu-4g-metal", "a4x-maxgpu-4g-metal-nolssd"])),
                     extra_spec={"tolerations": [{"operator": "Exists"}, {"key": "kubernetes.io/arch", "operator": "Equal", "value": "arm64", "effect": "NoSchedule"}]})


def tcpx_metrics_ds() -> dict:
    c = {"name": "tcpx-metrics", "image": IMG["tcpx-metrics"], "securityContext": {"privileged": True}, "env": env(PUSH_TO_CLOUD_MONITORING="true"),
         "volumeMounts": [mount("tcpx-socket", "/run/tcpx"), mount("library-dir-host", NVIDIA_CTR)]}
    return daemonset("tcpx-metrics-server", containers=[c], volumes=[host_vol("tcpx-socket", "/run/tcpx"), host_vol("library-dir-host", NVIDIA_HOST)],
                     affinity=node_affinity(expr(ACCEL_KEY, "In", ["nvidia-h100-80gb"])), critical=False)


# ---------------------------------------------------------------------------------------------------- nccl-test workloads
def headless_service(name: str) -> dict:
    return {"apiVersion": "v1", "kind": "Service", "metadata": {"name": name}, "spec": {"selector": {"name": name}, "clusterIP": "None"}}


def multi_nic_annotations(prefix: str, count: int, first_eth: int) -> dict:
    ifaces = [{"interfaceName": "eth0", "network": "default"}] + [{"interfaceName": f"eth{first_eth + i}", "network": f"{prefix}-{i}"} for i in range(count)]
    return {"networking.gke.io/default-interface": "eth0", "networking.gke.io/interfaces": json.dumps(ifaces)}


def nri_device_annotation(ctr: str, paths: list) -> dict:
    return {f"devices.gke.io/container.{ctr}": yaml.safe_dump([{"path": p} for p in paths], default_flow_style=False)}


GPU_DEV_PATHS = [f"/dev/nvidia{i}" for i in range(8)] + ["/dev/nvidiactl", "/dev/nvidia-uvm"]


def b200coll_test_pod(gpus: int = 8, autopilot: bool = False) -> list:
    """Single 8xB200 box: run the nccl-tests-style sweep of libb200coll next to stock NCCL (the repo's benchmark, BASELINE configs 2-4)."""
    cmd = ("source /usr/local/nvidia/lib64/b200coll-env-profile.sh\n"
           "for op in all_reduce all_gather reduce_scatter alltoall broadcast reduce; do\n"
           f"  /usr/local/nvidia/bin/${{op}}_perf --procs --ranks {gpus} -b 1K -e 1G -f 2 -w 5 --iters 100 -c 1 | tee /tmp/${{op}}_perf.txt\n"
           "done\nsleep infinity\n")
    pod = {"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "b200coll-test", "labels": {"name": "b200coll-test"}},
           "spec": {"hostNetwork": False, "hostPID": False, "restartPolicy": "Never",
                    "nodeSelector": {ACCEL_KEY: "nvidia-b200"},
                    "tolerations": [{"key": "nvidia.com/gpu", "operator": "Equal", "value": "present", "effect": "NoSchedule"}],
                    "volumes": [{"name": "shared-memory", "emptyDir": {"medium": "Memory", "sizeLimit": "250Gi"}}],
                    "containers": [{"name": "test", "image": IMG["b200coll"], "command": ["/bin/bash", "-c"], "args": [cmd],
                                    "resources": {"requests": {"cpu": "150m"}, "limits": {"nvidia.com/gpu": gpus}}, "volumeMounts": [mount("shared-memory", "/dev/shm")]}]}}
    if autopilot:
        pod["spec"]["nodeSelector"]["cloud.google.com/gke-gpu-driver-version"] = "latest"
    return [pod]


def nccl_pair(kind: str, variant: str) -> list:
    """Two pods + headless Services per transport, as in the reference's nccl-test*.yaml files."""
    docs = []
    for host in (1, 2):
        name = f"nccl-host-{host}"
        docs.append(headless_service(name))
    for host in (1, 2):
        docs.append(nccl_pod(kind, variant, host))
    return docs


def nccl_pod(kind: str, variant: str, host: int) -> dict:
    name = f"nccl-host-{host}"
    meta: dict = {"name": f"nccl-test-host-{host}", "labels": {"name": name}}
    spec: dict = {"hostNetwork": variant in ("base",) and kind in ("tcpx", "tcpxo"), "hostPID": False}
    vols = [host_vol("library-dir-host", NVIDIA_HOST), {"name": "shared-memory", "emptyDir": {"medium": "Memory", "sizeLimit": "250Gi"}}]
    test_mounts = [mount("library-dir-host", NVIDIA_CTR), mount("shared-memory", "/dev/shm")]
    test_env = env(LD_LIBRARY_PATH=f"{NVIDIA_CTR}/lib64")
    containers: list = []
    init: list = []
    ann: dict = {}
    gpus = 8
    if kind == "tcpx":
        rxdm_img = IMG["tcpx-rxdm-old"] if variant == "base" else IMG["tcpx-rxdm"]
        vols += [host_vol("tcpx-socket", "/run/tcpx"), {"name": "nccl-config", "configMap": {"name": "nccl-configmap", "defaultMode": 0o777}}, host_vol("sys", "/sys"), host_vol("proc-sys", "/proc/sys")]
        rxdm = {"name": "tcpx-daemon", "image": rxdm_img, "command": ["/tcpgpudmarxd/build/app/tcpgpudmarxd"], "args": ["--gpu_nic_preset", "a3vm", "--gpu_shmem_type", "fd", "--uds_path", "/run/tcpx", "--setup_param", "--verbose 128 2 0"],
                "env": env(LD_LIBRARY_PATH=f"{NVIDIA_CTR}/lib64"), "volumeMounts": [mount("library-dir-host", NVIDIA_CTR), mount("tcpx-socket", "/run/tcpx"), mount("sys", "/hostsysfs"), mount("proc-sys", "/hostprocsysfs")]}
        if variant in ("base", "without-hostnetwork"):
            rxdm["securityContext"] = {"privileged": True}
        else:
            rxdm["securityContext"] = {"capabilities": {"add": ["NET_ADMIN", "NET_BIND_SERVICE"]}}
            ann.update(nri_device_annotation("tcpx-daemon", GPU_DEV_PATHS))          # unprivileged sidecar gets the GPUs through the NRI injector
        if variant != "base":
            ann.update(multi_nic_annotations("vpc", 4, 1))
        test_mounts += [mount("tcpx-socket", "/tmp"), mount("nccl-config", "/configs")]
        test = {"name": "nccl-test", "image": IMG["tcpx-plugin"], "command": ["/bin/sh", "-c"], "args": ["/scripts/container_entry.sh shell\nsleep infinity\n"], "env": test_env, "volumeMounts": test_mounts,
                "resources": {"limits": {"nvidia.com/gpu": gpus}}}
        if variant == "base":
            test["securityContext"] = {"privileged": True}
        containers = [rxdm, test]
    elif kind == "tcpxo":
        rxdm_img = IMG["tcpxo-rxdm-old"] if variant == "base" else IMG["tcpxo-rxdm"]
        vols += [host_vol("aperture-devices", "/dev/aperture_devices"), host_vol("sys", "/sys"), host_vol("proc-sys", "/proc/sys")]
        rxdm = {"name": "tcpxo-daemon", "image": rxdm_img, "command": ["/bin/sh", "-c"],
                "args": ["set -ex\nchmod 755 /fts/entrypoint_rxdm_container.sh\n/fts/entrypoint_rxdm_container.sh --num_hops=2 --num_nics=8 --uid= --alsologtostderr\n"],
                "env": env(LD_LIBRARY_PATH=f"{NVIDIA_CTR}/lib64"), "volumeMounts": [mount("library-dir-host", NVIDIA_CTR), mount("sys", "/hostsysfs"), mount("proc-sys", "/hostprocsysfs")]}
        if variant == "base":
            rxdm["securityContext"] = {"privileged": True}
        else:
            rxdm["securityContext"] = {"capabilities": {"add": ["NET_ADMIN", "NET_BIND_SERVICE"]}}
            ann.update(nri_device_annotation("tcpxo-daemon", GPU_DEV_PATHS + ["/dev/dmabuf_import_helper"]))
            ann.update(multi_nic_annotations("vpc", 8, 1))
            spec["affinity"] = {"podAntiAffinity": {"requiredDuringSchedulingIgnoredDuringExecution": [{"labelSelector": {"matchExpressions": [expr("name", "In", ["nccl-host-1", "nccl-host-2"])]},
                                                                                                     "topologyKey": "kubernetes.io/hostname"}]}}
        test_mounts += [mount("aperture-devices", "/dev/aperture_devices")]
        test_env += env(NCCL_FASTRAK_LLCM_DEVICE_DIRECTORY="/dev/aperture_devices")
        args = "/scripts/container_entry.sh shell\nsource /usr/local/nvidia/lib64/nccl-env-profile.sh\nsleep infinity\n"
        if variant == "latest-autopilot":
            args = "/scripts/run-nccl-fastrak.sh all_gather_perf \"${LD_LIBRARY_PATH}\" 8 eth1,eth2,eth3,eth4,eth5,eth6,eth7,eth8 1M 1G 3 2 10\nsleep infinity\n"
        test = {"name": "nccl-test", "image": IMG["tcpxo-plugin"], "command": ["/bin/sh", "-c"], "args": [args], "env": test_env, "volumeMounts": test_mounts, "resources": {"limits": {"nvidia.com/gpu": gpus}}}
        if variant == "base":
            test["securityContext"] = {"privileged": True}
        if variant in ("latest", "latest-autopilot"):
            rxdm["restartPolicy"] = "Always"        # sidecar-style initContainer (gpudirect-tcpxo/best-practice.md:7-58)
            init = [rxdm]
            containers = [test]
        else:
            containers = [rxdm, test]
    else:   # rdma / gIB
        arm = variant in ("imex-a4x", "imex-a4x-max", "a4x-max")
        gpus = 4 if arm else 8
        image = {"imex-a4x": IMG["gib-diag-arm64"], "imex-a4x-max": IMG["gib-a4x-max"], "a4x-max": IMG["gib-a4x-max"]}.get(variant, IMG["gib-diag"])
        vols.append(host_vol("gib", "/home/kubernetes/bin/gib"))
        test_mounts.append(mount("gib", "/usr/local/gib"))
        if variant == "managed-rdma":
            spec["resourceClaims"] = [{"name": f"rdma-{i}", "resourceClaimTemplateName": "mrdma-claim"} for i in range(8)]
        elif variant == "a4x-max":
            spec["resourceClaims"] = [{"name": "compute-domain-channel", "resourceClaimTemplateName": "nccl-test-compute-domain-channel"}] + [{"name": f"rdma-{i}", "resourceClaimTemplateName": "mrdma-claim"} for i in range(4)]
        elif variant.startswith("imex"):
            spec["resourceClaims"] = [{"name": "compute-domain-channel", "resourceClaimTemplateName": "nccl-test-compute-domain-channel"}]
        else:
            ann.update(multi_nic_annotations("rdma", 8, 2))
        if arm:
            spec["nodeSelector"] = {"kubernetes.io/arch": "arm64"}
            spec["tolerations"] = [{"key": "kubernetes.io/arch", "operator": "Equal", "value": "arm64", "effect": "NoSchedule"}]
        if variant in ("a4-autopilot", "autopilot"):
            spec.setdefault("nodeSelector", {})["cloud.google.com/gke-gpu-driver-version"] = "latest"
        if variant in ("a4", "a4-autopilot"):
            spec.setdefault("nodeSelector", {})[ACCEL_KEY] = "nvidia-b200"
        elif variant in ("base", "autopilot", "managed-rdma"):
            spec.setdefault("nodeSelector", {})[ACCEL_KEY] = "nvidia-h200-141gb"
        test = {"name": "test", "image": image, "command": ["/bin/bash", "-c"], "args": ["/scripts/container_entry.sh shell\nsource /usr/local/gib/scripts/set_nccl_env.sh\nsleep infinity\n"],
                "env": test_env, "volumeMounts": test_mounts, "resources": {"requests": {"cpu": "150m"}, "limits": {"nvidia.com/gpu": gpus}}}
        if "resourceClaims" in spec:
            test["resources"]["claims"] = [{"name": c["name"]} for c in spec["resourceClaims"]]
        containers = [test]
    if ann:
        meta["annotations"] = ann
    spec["volumes"] = vols
    if init:
        spec["initContainers"] = init
    spec["containers"] = containers
    return {"apiVersion": "v1", "kind": "Pod", "metadata": meta, "spec": spec}


def rdma_extras(variant: str) -> list:
    docs = []
    if variant in ("imex-a4x", "imex-a4x-max", "a4x-max"):
        docs.append({"apiVersion": "resource.nvidia.com/v1beta1", "kind": "ComputeDomain", "metadata": {"name": "nccl-test-compute-domain"},
                     "spec": {"numNodes": 2, "channel": {"resourceClaimTemplate": {"name": "nccl-test-compute-domain-channel"}}}})
    if variant in ("managed-rdma", "a4x-max"):
        docs.append({"apiVersion": "resource.k8s.io/v1", "kind": "ResourceClaimTemplate", "metadata": {"name": "mrdma-claim"},
                     "spec": {"spec": {"devices": {"requests": [{"name": "rdma", "exactly": {"deviceClassName": "mrdma.google.com", "allocationMode": "ExactCount", "count": 1}}]}}}})
    return docs


def nccl_jobset() -> list:
    worker = {"name": "nccl-test", "image": IMG["gib-a4x-max"], "command": ["/bin/bash", "-c"], "args": [script("jobset-worker.sh")],
              "env": env(NUM_NODES="__NUM_NODES__", GPUS_PER_NODE=4, BENCHMARK="all_gather_perf", NCCL_TESTS_SPLIT_MASK="0x0", LD_LIBRARY_PATH=f"{NVIDIA_CTR}/lib64", REPLICATED_JOB_NAME="w") +
              [{"name": "JOBSET_NAME", "valueFrom": {"fieldRef": {"fieldPath": "metadata.annotations['jobset.sigs.k8s.io/jobset-name']"}}},
               {"name": "JOB_COMPLETION_INDEX", "valueFrom": {"fieldRef": {"fieldPath": "metadata.annotations['batch.kubernetes.io/job-completion-index']"}}}],
              "resources": {"limits": {"nvidia.com/gpu": 4}, "claims": [{"name": "compute-domain-channel"}]},
              "volumeMounts": [mount("library-dir-host", NVIDIA_CTR), mount("gib", "/usr/local/gib"), mount("shared-memory", "/dev/shm")]}
    pod_spec = {"restartPolicy": "Never", "nodeSelector": {"kubernetes.io/arch": "arm64"}, "subdomain": "nccl-ag",
                "tolerations": [{"key": "nvidia.com/gpu", "operator": "Equal", "value": "present", "effect": "NoSchedule"}, {"key": "kubernetes.io/arch", "operator": "Equal", "value": "arm64", "effect": "NoSchedule"}],
                "resourceClaims": [{"name": "compute-domain-channel", "resourceClaimTemplateName": "nccl-test-compute-domain-channel"}],
                "volumes": [host_vol("library-dir-host", NVIDIA_HOST), host_vol("gib", "/home/kubernetes/bin/gib"), {"name": "shared-memory", "emptyDir": {"medium": "Memory", "sizeLimit": "250Gi"}}],
                "containers": [worker]}
    js = {"apiVersion": "jobset.x-k8s.io/v1alpha2", "kind": "JobSet", "metadata": {"name": "nccl-ag"},
          "spec": {"ttlSecondsAfterFinished": 1200, "network": {"enableDNSHostnames": True},
                   "replicatedJobs": [{"name": "w", "replicas": 1, "template": {"spec": {"parallelism": "__NUM_NODES__", "completions": "__NUM_NODES__", "completionMode": "Indexed", "backoffLimit": 0,
                                                                                    "template": {"spec": pod_spec}}}}]}}
    return rdma_extras("imex-a4x-max") + [js]


def nccl_configmap() -> dict:
    return {"apiVersion": "v1", "kind": "ConfigMap", "metadata": {"name": "nccl-configmap"}, "data": {"allgather.sh": script("nccl-allgather.sh"), "run-nccl.sh": script("run-nccl.sh")}}


# ---------------------------------------------------------------------------------------------------- topology scheduler
def scheduler_docs(legacy: bool = False) -> dict:
    ns = "kube-system"
    sa = [{"apiVersion": "v1", "kind": "ServiceAccount", "metadata": {"name": "topology-scheduler", "namespace": ns}},
          {"apiVersion": "rbac.authorization.k8s.io/v1", "kind": "ClusterRole", "metadata": {"name": "topology-scheduler"},
           "rules": [{"apiGroups": [""], "resources": ["pods", "nodes", "namespaces"], "verbs": ["get", "list", "watch", "update", "patch"]}]},
          {"apiVersion": "rbac.authorization.k8s.io/v1", "kind": "ClusterRoleBinding", "metadata": {"name": "topology-scheduler"},
           "roleRef": {"apiGroup": "rbac.authorization.k8s.io", "kind": "ClusterRole", "name": "topology-scheduler"}, "subjects": [{"kind": "ServiceAccount", "name": "topology-scheduler", "namespace": ns}]}]
    args = ["python", "-m", "container_engine_accelerators_b200.scheduler.daemon", "--gate", "gke.io/topology-aware-auto-", "--interval", "1.0"] + (["--legacy-placement-group-key"] if legacy else [])
    dep = {"apiVersion": "apps/v1", "kind": "Deployment", "metadata": {"name": "topology-scheduler", "namespace": ns, "labels": {"app": "topology-scheduler"}},
           "spec": {"replicas": 1, "selector": {"matchLabels": {"app": "topology-scheduler"}},
                    "template": {"metadata": {"labels": {"app": "topology-scheduler"}},
                                 "spec": {"serviceAccountName": "topology-scheduler", "tolerations": [{"key": "components.gke.io/gke-managed-components", "operator": "Exists"}],
                                          "containers": [{"name": "topology-scheduler", "image": IMG["scheduler"], "command": args, "resources": {"requests": {"cpu": "100m", "memory": "128Mi"}}}]}}}}
    lab = {"name": "topology-labeler", "image": IMG["scheduler"], "command": ["python", "-m", "container_engine_accelerators_b200.scheduler.labeler", "--source", "gce-metadata"],
           "env": [{"name": "NODE_NAME", "valueFrom": {"fieldRef": {"fieldPath": "spec.nodeName"}}}], "resources": {"requests": {"cpu": "10m", "memory": "64Mi"}}}
    ds = daemonset("topology-labeler", containers=[lab], host_network=True, host_pid=False, critical=False, tolerate_all=False,
                   extra_spec={"serviceAccountName": "topology-scheduler", "tolerations": [{"key": "nvidia.com/gpu", "operator": "Equal", "value": "present", "effect": "NoSchedule"}]})
    return {"service-account.yaml": sa, "schedule-daemon.yaml": [dep], "label-nodes-daemon.yaml": [ds]}


# ---------------------------------------------------------------------------------------------------- demos / examples / tests
def demo_docs() -> dict:
    out: dict = {}
    serving = [{"apiVersion": "apps/v1", "kind": "Deployment", "metadata": {"name": "tf-serving", "labels": {"app": "tf-serving"}},
                "spec": {"replicas": 1, "selector": {"matchLabels": {"app": "tf-serving"}}, "template": {"metadata": {"labels": {"app": "tf-serving"}}, "spec": {
                    "containers": [{"name": "tf-serving", "image": "gcr.io/vishnuk-cloud/tf-serving:1.9-gpu-minimal", "ports": [{"containerPort": 8500}], "resources": {"limits": {"nvidia.com/gpu": 1}}}]}}}},
               {"apiVersion": "v1", "kind": "Service", "metadata": {"name": "tf-serving"}, "spec": {"selector": {"app": "tf-serving"}, "ports": [{"port": 8500, "targetPort": 8500}]}},
               {"apiVersion": "autoscaling/v2", "kind": "HorizontalPodAutoscaler", "metadata": {"name": "tf-serving"},
                "spec": {"scaleTargetRef": {"apiVersion": "apps/v1", "kind": "Deployment", "name": "tf-serving"}, "minReplicas": 1, "maxReplicas": 4,
                         "metrics": [{"type": "External", "external": {"metric": {"name": "kubernetes.io|container|accelerator|duty_cycle"}, "target": {"type": "AverageValue", "averageValue": "60"}}}]}}]
    out["demo/serving/tensorflow-serving.yaml"] = serving
    out["demo/serving/load_generator.yaml"] = [{"apiVersion": "apps/v1", "kind": "Deployment", "metadata": {"name": "load-generator"}, "spec": {"replicas": 1, "selector": {"matchLabels": {"app": "load-generator"}},
                                                "template": {"metadata": {"labels": {"app": "load-generator"}}, "spec": {"containers": [{"name": "client", "image": "gcr.io/vishnuk-cloud/tf-serving-client@sha256:869467cd5d22c14a024493327e323c69887f71476790a0e5762f2884b8e5773a",
                                                                                                                                    "args": ["--server=tf-serving:8500", "--concurrency=8"]}]}}}}]
    out["demo/image-prepull-ds.yaml"] = [daemonset("image-prepull", init=[{"name": "prepull", "image": IMG["b200coll"], "command": ["/bin/true"]}], host_network=False, host_pid=False, critical=False,
                                                   affinity=node_affinity(expr(ACCEL_KEY, "Exists")))]
    out["demo/device-plugin-health-monitoring-enabled.yaml"] = [device_plugin_ds(health=True, metrics=True)]
    for model in ("resnet", "inception-v3"):
        out[f"demo/tpu-training/{model}-tpu.yaml"] = [{"apiVersion": "batch/v1", "kind": "Job", "metadata": {"name": f"{model}-tpu"}, "spec": {"template": {"metadata": {"annotations": {"tf-version.cloud-tpus.google.com": "1.9"}},
                                                       "spec": {"restartPolicy": "Never", "containers": [{"name": model, "image": "gcr.io/tensorflow/tpu-models:r1.9", "command": ["python", f"/tensorflow_tpu_models/models/official/{model.split('-')[0]}/{model.replace('-', '_')}_main.py"],
                                                                                                             "resources": {"limits": {"cloud-tpus.google.com/v2": 8}}}]}}}}]
    # volume + claim in one file (the reference keeps demo/minikube/pv.yaml and pvc.yaml apart)
    out["demo/minikube/imagenet-storage.yaml"] = [
        {"apiVersion": "v1", "kind": "PersistentVolume", "metadata": {"name": "imagenet-pv", "labels": {"dataset": "imagenet"}},
         "spec": {"capacity": {"storage": "200Gi"}, "accessModes": ["ReadWriteOnce"], "persistentVolumeReclaimPolicy": "Retain", "hostPath": {"path": "/data/imagenet"}}},
        {"apiVersion": "v1", "kind": "PersistentVolumeClaim", "metadata": {"name": "imagenet-pvc"},
         "spec": {"accessModes": ["ReadWriteOnce"], "storageClassName": "", "selector": {"matchLabels": {"dataset": "imagenet"}}, "resources": {"requests": {"storage": "200Gi"}}}}]
    out["demo/minikube/resnet-gpu.yaml"] = [{"apiVersion": "batch/v1", "kind": "Job", "metadata": {"name": "resnet-gpu"}, "spec": {"template": {"spec": {"restartPolicy": "Never", "volumes": [{"name": "data", "persistentVolumeClaim": {"claimName": "imagenet-pvc"}}],
                                                                                                                                                      "containers": [{"name": "resnet", "image": "gcr.io/vishnuk-cloud/tf-models-gpu:1.0", "volumeMounts": [mount("data", "/data")], "resources": {"limits": {"nvidia.com/gpu": 1}}}]}}}}]
    out["demo/gpu-error/xid-inject-job.yaml"] = [{"apiVersion": "batch/v1", "kind": "Job", "metadata": {"name": "xid-inject"}, "spec": {"backoffLimit": 0, "template": {"spec": {"restartPolicy": "Never",
                                                 "containers": [{"name": "xid-inject", "image": IMG["xid-inject"], "command": ["/usr/bin/xid_inject", "--mode", "oob-store"], "resources": {"limits": {"nvidia.com/gpu": 1}}}]}}}}]
    out["example/tensorflow-notebook/tensorflow-notebook.yaml"] = [{"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "tensorflow-notebook", "labels": {"app": "notebook"}},
                                                                    "spec": {"containers": [{"name": "notebook", "image": "gcr.io/kubeflow/tensorflow-notebook-cpu:v1", "ports": [{"containerPort": 8888}], "resources": {"limits": {"nvidia.com/gpu": 1}}}]}}]
    out["example/cuda-mps/mps-probe.yaml"] = [{"apiVersion": "v1", "kind": "Pod", "metadata": {"name": "mps-probe"}, "spec": {"restartPolicy": "Never", "hostIPC": True,
                                               "containers": [{"name": "probe", "image": IMG["b200coll"], "command": ["/usr/local/nvidia/bin/mps_probe", "--json"], "resources": {"limits": {"nvidia.com/gpu": 1}}}]}}]
    return out


def test_fixture_docs() -> dict:
    sel = "cloud.google.com/gke-accelerator-test"
    mig = cos_driver_ds("nvidia-mig"); pre = cos_driver_ds("preloaded"); ub = ubuntu_driver_ds(pin="470.103.01")
    for ds in (mig, pre, ub):
        ds["spec"]["template"]["spec"]["affinity"] = node_affinity(expr("gke-accelerator-test", "Exists"))
    return {"test/nvidia_gpu/device-plugin-test.yaml": [device_plugin_ds(name="b200-gpu-device-plugin-test", mps=True, selector_key="gke-accelerator-test")] + device_plugin_rbac(),
            "test/nvidia_gpu/daemonset-nvidia-mig-test.yaml": [mig], "test/nvidia_gpu/daemonset-nvidia-preloaded-test.yaml": [pre], "test/nvidia_gpu/daemonset-ubuntu-preloaded.yaml": [ub],
            "test/nvidia_gpu/xid-config.yaml": [xid_config("32,79,74")]}


# ---------------------------------------------------------------------------------------------------- the tree
def build_tree() -> dict:
    t: dict = {}
    t["device-plugin/device-plugin.yaml"] = [device_plugin_ds()]
    t["device-plugin/device-plugin-native.yaml"] = [device_plugin_native_ds()]
    t["device-plugin/rbac.yaml"] = device_plugin_rbac()
    t["device-plugin/xid-config.yaml"] = [xid_config()]
    t["device-plugin/gpu-config-b200coll.yaml"] = [gpu_config_map()]
    t["device-plugin/gpu-config-mig-1g23gb.yaml"] = [gpu_config_map(partition="1g.23gb")]
    t["partition-gpu/partition-gpu.yaml"] = [partition_gpu_ds()]
    t["nri-device-injector/nri-device-injector.yaml"] = [nri_injector_ds()]
    t["nri-device-injector/nri-device-injector-autopilot.yaml"] = [{"apiVersion": "v1", "kind": "Namespace", "metadata": {"name": "gpudirect-system"}}, nri_injector_ds(autopilot=True)]
    cos = ["preloaded", "preloaded-latest", "preloaded-latest-a4x", "confidential", "confidential-latest", "nvidia-mig", "vgpu-latest"]
    for v in cos:
        t[f"driver-installer/cos/daemonset-{v}.yaml"] = [cos_driver_ds(v)]
    t["driver-installer/cos/kustomization.yaml"] = [{"apiVersion": "kustomize.config.k8s.io/v1beta1", "kind": "Kustomization", "resources": ["daemonset-preloaded.yaml"]}]
    # one-command install of the node side for an 8xB200 (a4-highgpu-8g) pool: `kubectl apply -k deploy/overlays/a4-b200`
    t["overlays/a4-b200/kustomization.yaml"] = [{"apiVersion": "kustomize.config.k8s.io/v1beta1", "kind": "Kustomization", "namespace": "kube-system",
                                                 "resources": ["../../device-plugin/rbac.yaml", "../../device-plugin/xid-config.yaml", "../../device-plugin/gpu-config-b200coll.yaml",
                                                               "../../device-plugin/device-plugin-native.yaml", "../../transport/b200coll-installer.yaml",
                                                               "../../nri-device-injector/nri-device-injector.yaml"]}]
    t["overlays/a4-b200-mig/kustomization.yaml"] = [{"apiVersion": "kustomize.config.k8s.io/v1beta1", "kind": "Kustomization", "namespace": "kube-system",
                                                     "resources": ["../../device-plugin/rbac.yaml", "../../device-plugin/xid-config.yaml", "../../device-plugin/gpu-config-mig-1g23gb.yaml",
                                                                   "../../partition-gpu/partition-gpu.yaml", "../../device-plugin/device-plugin-native.yaml"]}]
    t["driver-installer/ubuntu/daemonset.yaml"] = [ubuntu_driver_ds(preloaded=False)]
    t["driver-installer/ubuntu/daemonset-preloaded.yaml"] = [ubuntu_driver_ds()]
    for rel, ver in (("R525", "525.147.05"), ("R535", "535.230.02"), ("R550", "550.144.03"), ("R570", "570.124.06")):
        t[f"driver-installer/ubuntu/daemonset-preloaded-{rel}.yaml"] = [ubuntu_driver_ds(pin=ver)]
    t["driver-installer/minikube/daemonset.yaml"] = [minikube_driver_ds()]
    t["driver-installer/daemonset.yaml"] = [legacy_root_ds()]
    t["transport/b200coll-installer.yaml"] = [b200coll_installer_ds()]
    t["transport/b200coll-installer-autopilot.yaml"] = [{"apiVersion": "v1", "kind": "Namespace", "metadata": {"name": "b200coll-system"}}, b200coll_installer_ds(autopilot=True)]
    t["transport/compat/fast-socket-installer.yaml"] = [compat_installer_ds("fast-socket")]
    for kind in ("tcpx", "tcpxo"):
        t[f"transport/compat/nccl-{kind}-installer.yaml"] = [compat_installer_ds(kind)]
        t[f"transport/compat/nccl-{kind}-installer-autopilot.yaml"] = [compat_installer_ds(kind, autopilot=True)]
    t["transport/compat/nccl-rdma-installer.yaml"] = [compat_installer_ds("rdma")]
    t["transport/compat/nccl-rdma-installer-a4x.yaml"] = [compat_installer_ds("rdma", arm64=True)]
    t["transport/compat/nccl-rdma-installer-autopilot.yaml"] = [compat_installer_ds("rdma", autopilot=True)]
    t["transport/compat/asapd-lite-installer-a4x-max-bm-cos.yaml"] = [asapd_lite_ds()]
    t["transport/compat/optmem-max-ds.yaml"] = [host_tweak_ds("optmem-max", "optmem-max.sh", accel=["nvidia-h100-80gb"], nsenter=False)]
    t["transport/compat/tcpx-metrics-server.yaml"] = [tcpx_metrics_ds()]
    t["transport/compat/cos-enable-kdump.yaml"] = [host_tweak_ds("cos-enable-kdump", "cos-enable-kdump.sh", extra_affinity=[expr("cloud.google.com/gke-os-distribution", "In", ["cos"]), expr("gke-kdump-enabled", "In", ["true"])])]
    t["transport/compat/fix-hostname.yaml"] = [host_tweak_ds("fix-hostname", "fix-hostname.sh")]
    t["nccl-test/b200coll-test.yaml"] = b200coll_test_pod()
    t["nccl-test/b200coll-test-autopilot.yaml"] = b200coll_test_pod(autopilot=True)
    t["nccl-test/tcpx/nccl-config.yaml"] = [nccl_configmap()]
    for v, fn in (("base", "nccl-test"), ("latest", "nccl-test-latest"), ("latest-autopilot", "nccl-test-latest-autopilot"), ("unprivileged-without-hostnetwork", "nccl-test-unprivileged-without-hostnetwork"),
                  ("without-hostnetwork", "nccl-test-without-hostnetwork")):
        t[f"nccl-test/tcpx/{fn}.yaml"] = nccl_pair("tcpx", v)
    for v, fn in (("base", "nccl-test"), ("latest", "nccl-test-latest"), ("latest-autopilot", "nccl-test-latest-autopilot"), ("unprivileged-without-hostnetwork", "nccl-test-unprivileged-without-hostnetwork")):
        t[f"nccl-test/tcpxo/{fn}.yaml"] = nccl_pair("tcpxo", v)
    for v, fn in (("base", "nccl-test"), ("a4", "nccl-test-a4"), ("a4-autopilot", "nccl-test-a4-autopilot"), ("autopilot", "nccl-test-autopilot"), ("managed-rdma", "nccl-test-managed-rdma"),
                  ("imex-a4x", "nccl-test-imex-a4x"), ("imex-a4x-max", "nccl-test-imex-a4x-max"), ("a4x-max", "nccl-test-a4x-max")):
        t[f"nccl-test/rdma/{fn}.yaml"] = rdma_extras(v) + nccl_pair("rdma", v)
    t["nccl-test/rdma/nccl-test-a4x-max-jobset.yaml"] = nccl_jobset()
    for legacy, d in ((False, "topology-scheduler"), (True, "topology-scheduler/legacy-tcpxo")):
        for fn, docs in scheduler_docs(legacy).items():
            t[f"{d}/{fn}"] = docs
    t.update(demo_docs())
    t.update(test_fixture_docs())
    return t


HEADER = "# GENERATED by deploy/generate.py — edit the generator, not this file.\n"


class _Dumper(yaml.SafeDumper):
    pass


def _str_representer(dumper, data):
    return dumper.represent_scalar("tag:yaml.org,2002:str", data, style="|" if "\n" in data else None)


_Dumper.add_representer(str, _str_representer)


def render(docs: list) -> str:
    return HEADER + "---\n".join(yaml.dump(d, Dumper=_Dumper, sort_keys=True, default_flow_style=False, width=200) for d in docs)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true", help="exit 1 if any generated file is missing or stale")
    args = ap.parse_args(argv)
    stale = []
    for rel, docs in sorted(build_tree().items()):
        path = HERE / rel
        text = render(docs)
        if args.check:
            if not path.exists() or path.read_text() != text:
                stale.append(rel)
        else:
            path.parent.mkdir(parents=True, exist_ok=True)
            path.write_text(text)
    if args.check and stale:
        print("stale or missing manifests (run python deploy/generate.py):\n  " + "\n  ".join(stale))
        return 1
    print(f"{'checked' if args.check else 'wrote'} {len(build_tree())} manifests")
    return 0


if __name__ == "__main__":
    sys.exit(main())
