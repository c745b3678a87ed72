"""GPU manager: discovery, virtual-device fan-out, DeviceSpec building, MPS env maths, and the
serve / re-register / re-discover loop around the DevicePlugin gRPC server.

Contract: reference pkg/gpu/nvidia/manager.go:140-550. Deliberate fixes of reference bugs
(SURVEY §7.3-9): every read of the device map holds the lock (manager.go:173-232 does not); the health queue
is bounded and never blocks the NVML listener (health_checker.go:421,442); the kubelet-socket watcher error is
not ignored (manager.go:444-445); re-discovery failures back off instead of spinning (manager.go:516-521).
"""
from __future__ import annotations

import logging
import os
import queue
import subprocess
import threading
import time
from concurrent import futures
from dataclasses import dataclass
from typing import Optional

import grpc

from . import mig as migmod
from . import nvml as nvmlmod
from . import protos, sharing
from .config import GPUConfig

log = logging.getLogger("b200-device-plugin")

RESOURCE_NAME = "nvidia.com/gpu"
NVIDIA_CTL, NVIDIA_UVM, NVIDIA_UVM_TOOLS, NVIDIA_MODESET = "nvidiactl", "nvidia-uvm", "nvidia-uvm-tools", "nvidia-modeset"
GPU_CHECK_INTERVAL = 10.0
PLUGIN_SOCKET_CHECK_INTERVAL = 1.0
MPS_DIR = "/tmp/nvidia-mps"
MPS_CONTROL_BIN = "/usr/local/nvidia/bin/nvidia-cuda-mps-control"
MPS_ACTIVE_THREAD_CMD = "get_default_active_thread_percentage"
MPS_MEM_LIMIT_ENV = "CUDA_MPS_PINNED_DEVICE_MEM_LIMIT"
MPS_THREAD_LIMIT_ENV = "CUDA_MPS_ACTIVE_THREAD_PERCENTAGE"


@dataclass
class Mount:
    host_path: str
    container_path: str
    read_only: bool = True


class AllocationError(ValueError):
    pass


class GPUManager:
    def __init__(self, dev_dir: str, proc_dir: str, mount_paths: list, gpu_config: GPUConfig, nvml: Optional[nvmlmod.NvmlOperations] = None,
                 pci_root: str = nvmlmod.PCI_DEVICES_ROOT, mps_control_bin: str = MPS_CONTROL_BIN,
                 gpu_check_interval: float = GPU_CHECK_INTERVAL, socket_check_interval: float = PLUGIN_SOCKET_CHECK_INTERVAL,
                 preferred_allocation_policy: str = "none"):
        self.dev_dir, self.proc_dir = dev_dir, proc_dir
        self.preferred_allocation_policy = preferred_allocation_policy     # "none" = the reference's contract (no plugin options)
        self.mount_paths: list = list(mount_paths)
        self.gpu_config = gpu_config
        self.nvml = nvml
        self.pci_root = pci_root
        self.mps_control_bin = mps_control_bin
        self.gpu_check_interval, self.socket_check_interval = gpu_check_interval, socket_check_interval
        self.default_devices: list = []
        self.devices: dict = {}                       # name -> migmod.Device
        self.lock = threading.RLock()
        self.ctl_path = os.path.join(dev_dir, NVIDIA_CTL)
        self.uvm_path = os.path.join(dev_dir, NVIDIA_UVM)
        self.mig = migmod.MigDeviceManager(dev_dir, proc_dir, topology_of=self._topology_by_index)
        self.health: "queue.Queue[migmod.Device]" = queue.Queue(maxsize=1024)
        self.total_mem_per_gpu = 0
        self.grpc_server: Optional[grpc.Server] = None
        self.socket_path = ""
        self._stop = threading.Event()
        self._restart = threading.Event()
        self._watchers: list = []                     # ListAndWatch streams to wake on stop
        self.serving = threading.Event()

    # ------------------------------------------------------------------ inventory
    def _topology_by_index(self, index: int) -> Optional[int]:
        if self.nvml is None:
            return None
        info = self.nvml.device(index)
        return nvmlmod.numa_topology(info.bus_id, self.pci_root) if info.bus_id else None

    def list_physical_devices(self) -> dict:
        with self.lock:
            if not self.gpu_config.gpu_partition_size:
                return dict(self.devices)
            return dict(self.mig.list_partitions())

    def list_health_critical_xid(self) -> list:
        return list(self.gpu_config.health_critical_xid)

    def list_devices(self) -> dict:
        """Physical devices, or <id>/vgpu<k> fan-out under sharing; vGPUs inherit the parent's health."""
        physical = self.list_physical_devices()
        n = self.gpu_config.sharing.max_shared_clients_per_gpu
        if n > 0:
            out = {}
            for dev in physical.values():
                for i in range(n):
                    vid = f"{dev.id}/vgpu{i}"
                    out[vid] = migmod.Device(vid, dev.health, dev.numa_node)
            return out
        return physical

    def device_spec(self, device_id: str) -> list:
        if self.gpu_config.sharing.max_shared_clients_per_gpu > 0:
            device_id = sharing.virtual_to_physical_device_id(device_id)
        with self.lock:
            if not self.gpu_config.gpu_partition_size:
                dev = self.devices.get(device_id)
                if dev is None:
                    raise AllocationError(f"invalid allocation request with non-existing device {device_id}")
                if dev.health != protos.HEALTHY:
                    raise AllocationError(f"invalid allocation request with unhealthy device {device_id}")
                p = os.path.join(self.dev_dir, device_id)
                return [migmod.Spec(p, p)]
            try:
                return list(self.mig.device_spec(device_id))
            except migmod.MigError as e:
                raise AllocationError(str(e)) from e

    def set_device_health(self, name: str, health: str, numa_node: Optional[int] = None) -> None:
        with self.lock:
            if nvmlmod.NVIDIA_DEVICE_RE.match(name):
                self.devices[name] = migmod.Device(name, health, numa_node)
            else:
                self.mig.set_device_health(name, health, numa_node)

    def report_unhealthy(self, dev: migmod.Device) -> bool:
        """Called by the health checker. Never blocks the NVML listener: a full queue drops (and logs)."""
        try:
            self.health.put_nowait(dev)
            return True
        except queue.Full:
            log.error("health queue full; dropping update for %s", dev.id)
            return False

    # ------------------------------------------------------------------ discovery
    def discover_gpus(self) -> None:
        if self.nvml is None:
            self.nvml = nvmlmod.NativeNvml()
        count = self.nvml.device_count()
        for i in range(count):
            info = self.nvml.device(i)
            name = f"nvidia{info.minor}"
            numa = None
            try:
                numa = nvmlmod.numa_topology(info.bus_id, self.pci_root) if info.bus_id else None
            except nvmlmod.NvmlError as e:
                log.error("unable to get topology for device with index %d: %s", i, e)
            self.set_device_health(name, protos.HEALTHY, numa)

    def discover_num_gpus(self) -> int:
        return sum(1 for e in os.scandir(self.dev_dir) if not e.is_dir() and nvmlmod.NVIDIA_DEVICE_RE.match(e.name))

    def has_additional_gpus_installed(self) -> bool:
        with self.lock:
            original = len(self.devices)
        try:
            count = self.discover_num_gpus()
        except OSError as e:
            log.error("%s", e)
            return False
        if count > original:
            log.info("Found %d GPUs, while only %d are registered. Stopping device-plugin server.", count, original)
            return True
        return False

    def check_device_paths(self) -> None:
        os.stat(self.ctl_path)
        os.stat(self.uvm_path)

    def is_mps_healthy(self) -> None:
        try:
            proc = subprocess.run([self.mps_control_bin], input=MPS_ACTIVE_THREAD_CMD.encode(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=30)
        except (OSError, subprocess.TimeoutExpired) as e:
            raise RuntimeError(f"failed to start NVIDIA MPS health check command: {e}") from e
        if proc.returncode != 0:
            raise RuntimeError(f"failed to health check NVIDIA MPS: exit status {proc.returncode}")
        log.info("MPS is healthy, active thread percentage = %s", proc.stdout.decode().strip())

    def envs(self, num_devices_requested: int) -> dict:
        env = {}
        sh = self.gpu_config.sharing
        if sh.strategy == sharing.MPS:
            active = num_devices_requested * 100 // sh.max_shared_clients_per_gpu
            mem = num_devices_requested * self.total_mem_per_gpu // sh.max_shared_clients_per_gpu
            env = {MPS_THREAD_LIMIT_ENV: str(active), MPS_MEM_LIMIT_ENV: f"0={mem // (1024 * 1024)}M"}
        return env

    def start(self) -> None:
        self.default_devices = [self.ctl_path, self.uvm_path]
        for extra in (NVIDIA_MODESET, NVIDIA_UVM_TOOLS):
            p = os.path.join(self.dev_dir, extra)
            if os.path.exists(p):
                self.default_devices.append(p)
        self.discover_gpus()
        if self.gpu_config.gpu_partition_size:
            try:
                with self.lock:
                    self.mig.start(self.gpu_config.gpu_partition_size)
            except migmod.MigError as e:
                raise RuntimeError(f"failed to start mig device manager: {e}") from e
        if self.gpu_config.sharing.strategy == sharing.MPS:
            try:
                self.is_mps_healthy()
            except RuntimeError as e:
                raise RuntimeError(f"NVIDIA MPS is not running on this node: {e}") from e
            if not any(m.host_path == MPS_DIR for m in self.mount_paths):
                self.mount_paths.append(Mount(MPS_DIR, MPS_DIR, read_only=False))
            if self.nvml.device_count() <= 0:
                raise RuntimeError("failed to query total memory available per GPU: no GPUs on node")
            self.total_mem_per_gpu = self.nvml.device(0).mem_total

    # ------------------------------------------------------------------ serving
    def stop(self) -> None:
        self._stop.set()
        self._restart.set()

    def _stop_server(self) -> None:
        srv, self.grpc_server = self.grpc_server, None
        self.serving.clear()
        if srv is not None:
            srv.stop(grace=0.2)
        try:
            if self.socket_path and os.path.exists(self.socket_path):
                os.unlink(self.socket_path)
        except OSError:
            pass

    def serve(self, plugin_dir: str, kubelet_endpoint: str, plugin_endpoint: str, max_restarts: Optional[int] = None) -> None:
        """Blocks. Restart triggers (reference manager.go:501-534): plugin socket vanished (1 s poll), more
        /dev/nvidiaN than registered (10 s poll -> rediscover), kubelet.sock re-created (kubelet restarted)."""
        from .plugin import DevicePluginService, register_with_kubelet
        kubelet_path = os.path.join(plugin_dir, kubelet_endpoint)
        register = os.path.exists(kubelet_path)
        log.info("registered with kubelet, will use beta API" if register else "no kubelet.sock to register.")
        restarts = 0
        while not self._stop.is_set():
            self._restart.clear()
            self.socket_path = os.path.join(plugin_dir, plugin_endpoint)
            try:
                if os.path.exists(self.socket_path):
                    os.unlink(self.socket_path)
            except OSError as e:
                log.error("cannot remove stale socket %s: %s", self.socket_path, e)
            server = grpc.server(futures.ThreadPoolExecutor(max_workers=8), options=[("grpc.so_reuseport", 0)])
            service = DevicePluginService(self)
            service.register(server)
            server.add_insecure_port(f"unix:{self.socket_path}")
            server.start()
            self.grpc_server = server
            log.info("device-plugin: serving on %s", self.socket_path)
            # identity of the kubelet socket BEFORE registering: a kubelet that restarts right after our Register call then shows up as a
            # different socket in the watch loop below (taken afterwards, the new kubelet would silently become the baseline)
            kubelet_ino = self._ino(kubelet_path)
            if register:
                try:
                    register_with_kubelet(kubelet_path, plugin_endpoint, RESOURCE_NAME, preferred_allocation=self.preferred_allocation_policy != "none")
                    log.info("device-plugin registered with the kubelet")
                except Exception as e:
                    self._stop_server()
                    raise RuntimeError(f"device-plugin: cannot register to kubelet service: {e}") from e
            self.serving.set()
            kubelet_gone = False
            next_gpu_check = time.monotonic() + self.gpu_check_interval
            while not self._restart.is_set():
                self._restart.wait(self.socket_check_interval)
                if self._stop.is_set():
                    break
                if not os.path.lexists(self.socket_path):
                    log.info("plugin socket %s was removed; restarting the server", self.socket_path)
                    break
                ino = self._ino(kubelet_path)
                if register and ino is None:
                    kubelet_gone = True          # seen while the socket was absent: whatever appears next is a new kubelet
                if register and ino is not None and (ino != kubelet_ino or kubelet_gone):
                    log.info("kubelet socket was re-created (kubelet restart); re-registering")
                    break
                if not register and ino is not None:
                    register = True
                    log.info("kubelet socket appeared; registering")
                    break
                if time.monotonic() >= next_gpu_check:
                    next_gpu_check = time.monotonic() + self.gpu_check_interval
                    if self.has_additional_gpus_installed():
                        self._stop_server()
                        backoff = 1.0
                        while not self._stop.is_set():
                            try:
                                self.discover_gpus()
                                break
                            except Exception as e:     # back off instead of spinning (reference manager.go:516-521 spins)
                                log.error("rediscovery failed: %s; retrying in %.0fs", e, backoff)
                                self._stop.wait(backoff)
                                backoff = min(backoff * 2, 30.0)
                        break
            self._stop_server()
            restarts += 1
            if max_restarts is not None and restarts >= max_restarts:
                break

    @staticmethod
    def _ino(path: str):
        """Identity of a socket file: (inode, change time). The inode number alone is not enough: a kubelet that removes and re-creates
        kubelet.sock between two polls usually gets the same number back from the filesystem."""
        try:
            st = os.stat(path)
            return (st.st_ino, st.st_ctime_ns)
        except OSError:
            return None
