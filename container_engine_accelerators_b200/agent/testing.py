"""Conformance fixtures: a stub kubelet (Registration gRPC server on <dir>/kubelet.sock) and builders for fake
/dev, /proc and /sys trees — the reference's test techniques (pkg/gpu/nvidia/beta_plugin_test.go:36-70 KubeletStub;
pkg/gpu/nvidia/mig/mig_test.go:55-83 fake capability files; manager_test.go:223-310 fake numa_node files) as a reusable,
language-neutral harness: anything that speaks v1beta1 over a Unix socket can be driven by it.
"""
from __future__ import annotations

import os
import threading
from concurrent import futures
from pathlib import Path

import grpc

from . import protos
from .protos import deviceplugin as pb


class KubeletStub:
    """Records every RegisterRequest; `wait_registration()` blocks until one arrives."""

    def __init__(self, plugin_dir: str, socket_name: str = "kubelet.sock"):
        self.socket_path = os.path.join(plugin_dir, socket_name)
        self.requests: list = []
        self._event = threading.Event()
        self._server = None

    def _register(self, request, context):
        self.requests.append(request)
        self._event.set()
        return pb.Empty()

    def start(self) -> "KubeletStub":
        if os.path.exists(self.socket_path):
            os.unlink(self.socket_path)
        self._server = grpc.server(futures.ThreadPoolExecutor(max_workers=2))
        handler = grpc.method_handlers_generic_handler(protos.REGISTRATION_SERVICE, {
            "Register": grpc.unary_unary_rpc_method_handler(self._register, pb.RegisterRequest.FromString, pb.Empty.SerializeToString)})
        self._server.add_generic_rpc_handlers((handler,))
        self._server.add_insecure_port(f"unix:{self.socket_path}")
        self._server.start()
        return self

    def wait_registration(self, timeout: float = 10.0):
        if not self._event.wait(timeout):
            raise TimeoutError("plugin did not register with the stub kubelet")
        self._event.clear()
        return self.requests[-1]

    def stop(self) -> None:
        if self._server:
            self._server.stop(grace=0)
            self._server = None
        if os.path.exists(self.socket_path):
            os.unlink(self.socket_path)


def make_fake_dev(root: str, gpus: int = 2, with_optional: bool = True) -> str:
    """<root>/dev with nvidiactl, nvidia-uvm, [nvidia-uvm-tools, nvidia-modeset] and nvidia0..N-1 as empty files."""
    dev = Path(root) / "dev"
    dev.mkdir(parents=True, exist_ok=True)
    names = ["nvidiactl", "nvidia-uvm"] + (["nvidia-uvm-tools", "nvidia-modeset"] if with_optional else [])
    for n in names + [f"nvidia{i}" for i in range(gpus)]:
        (dev / n).touch()
    return str(dev)


def add_fake_gpu(dev_dir: str, index: int) -> None:
    (Path(dev_dir) / f"nvidia{index}").touch()


def make_fake_mig(root: str, dev_dir: str, gpus: int, partitions_per_gpu: int) -> str:
    """<root>/proc/driver/nvidia/capabilities/gpu<N>/mig/gi<M>/{access,ci0/access} + /dev/nvidia-caps/nvidia-cap<minor>."""
    proc = Path(root) / "proc"
    caps = Path(dev_dir) / "nvidia-caps"
    caps.mkdir(parents=True, exist_ok=True)
    minor = 10
    for g in range(gpus):
        base = proc / "driver/nvidia/capabilities" / f"gpu{g}" / "mig"
        for p in range(partitions_per_gpu):
            gi = base / f"gi{p + 1}"
            (gi / "ci0").mkdir(parents=True, exist_ok=True)
            (gi / "access").write_text(f"DeviceFileMinor: {minor}\nDeviceFileMode: 292\n")
            (caps / f"nvidia-cap{minor}").touch(); minor += 1
            (gi / "ci0" / "access").write_text(f"DeviceFileMinor: {minor}\nDeviceFileMode: 292\n")
            (caps / f"nvidia-cap{minor}").touch(); minor += 1
    (proc / "driver/nvidia/capabilities" / "mig").mkdir(parents=True, exist_ok=True)   # non-gpu entry must be skipped
    return str(proc)


def make_fake_pci(root: str, bus_id: str, numa_node: int) -> str:
    pci = Path(root) / "sys/bus/pci/devices"
    d = pci / bus_id
    d.mkdir(parents=True, exist_ok=True)
    (d / "numa_node").write_text(f"{numa_node}\n")
    return str(pci)
