"""Prometheus metrics server: per-container and per-node GPU duty cycle / memory gauges + request counts,
and (new) per-collective counters read from the stats pages libb200coll exports in /dev/shm.

Contract: reference pkg/gpu/nvidia/metrics/{metrics,devices,util}.go: gauge names and labels,
`make="nvidia"`, accelerator_id = GPU UUID, model = NVML name, resource_name = nvidia.com/gpu; duty cycle =
integer mean of utilisation samples newer than now-10 s, > 100 => device skipped this tick; container->device map
from the kubelet PodResources v1alpha1 List, virtual (shared) ids dropped; all series reset once a minute so dead
containers disappear. Fixes: connection closed only if it was opened (devices.go:59-68); zero-sample guard lives
in the native sampler (util.go:82).
"""
from __future__ import annotations

import glob
import logging
import struct
import threading
import time
from typing import Optional

import grpc
from prometheus_client import CollectorRegistry, Gauge, start_http_server

from . import protos, sharing
from .protos import podresources as prpb

log = logging.getLogger("b200-device-plugin")

GPU_RESOURCE_NAME = "nvidia.com/gpu"
POD_RESOURCES_SOCKET = "/var/lib/kubelet/pod-resources/kubelet.sock"
RESET_INTERVAL_S = 60.0
DUTY_CYCLE_WINDOW_S = 10
_CONTAINER_LABELS = ["namespace", "pod", "container", "make", "accelerator_id", "model"]
_NODE_LABELS = ["make", "accelerator_id", "model"]
COLL_OPS = ("all_reduce", "all_gather", "reduce_scatter", "alltoall", "broadcast", "reduce")
COLL_ALGOS = ("auto", "ll", "oneshot", "twoshot", "nvls", "copy", "ll2")


def get_devices_for_all_containers(socket_path: str = POD_RESOURCES_SOCKET, timeout: float = 5.0) -> dict:
    """{(namespace, pod, container): [device ids]} from the kubelet's PodResourcesLister."""
    out: dict = {}
    try:
        channel = grpc.insecure_channel(f"unix:{socket_path}")
    except Exception as e:
        raise RuntimeError(f"error connecting to kubelet PodResourceLister service: {e}") from e
    try:
        # The reference speaks v1alpha1 (metrics/devices.go:33-34). Kubelets since 1.20 also serve v1, whose messages are a field-number
        # compatible superset, and may stop serving v1alpha1: ask for v1 first and fall back when the kubelet does not know it.
        resp, last = None, None
        for service in (protos.POD_RESOURCES_SERVICE_V1, protos.POD_RESOURCES_SERVICE):
            call = channel.unary_unary(f"/{service}/List", request_serializer=prpb.ListPodResourcesRequest.SerializeToString,
                                       response_deserializer=prpb.ListPodResourcesResponse.FromString)
            try:
                resp = call(prpb.ListPodResourcesRequest(), timeout=timeout)
                break
            except grpc.RpcError as e:
                last = e
                if e.code() != grpc.StatusCode.UNIMPLEMENTED:
                    break
        if resp is None:
            raise RuntimeError(f"error listing pod resources: {last}") from last
        for pod in resp.pod_resources:
            for c in pod.containers:
                key = (pod.namespace, pod.name, c.name)
                for d in c.devices:
                    if not d.device_ids or d.resource_name != GPU_RESOURCE_NAME:
                        continue
                    ids = [i for i in d.device_ids if not sharing.is_virtual_device_id(i)]
                    out.setdefault(key, []).extend(ids)
    finally:
        channel.close()
    return out


COLL_STATS_STALE_S = 3600      # pages of processes that died without CommDestroy stop being exported after this long without an update


def read_coll_stats_pages(pattern: str = "/dev/shm/b200coll.*", now=time.time) -> list:
    """Parse the 4 KiB pages written by libb200coll (coll/src/comm.cu stats_page_publish). The library refreshes its page every 256
    collective calls; a page whose header timestamp is older than COLL_STATS_STALE_S belongs to a dead or idle process and is skipped
    (the exporter runs in its own PID namespace, so it cannot ask whether the pid is alive)."""
    pages = []
    for path in glob.glob(pattern):
        try:
            with open(path, "rb") as f:
                raw = f.read(4096)
            if len(raw) < 64 + 8 * 21 or raw[:8] != b"B200COLL":
                continue
            version, pid, rank, nranks, device, nvls = struct.unpack_from("<6I", raw, 8)
            nops = 4 if version == 1 else 6          # v1 pages (older library): no broadcast / reduce counters
            updated = struct.unpack_from("<Q", raw, 32)[0] if version >= 2 else 0
            if updated and now() - updated > COLL_STATS_STALE_S:
                continue
            vals = struct.unpack_from(f"<{2 * nops + 9}Q", raw, 64)
            p2p = struct.unpack_from("<3Q", raw, 64 + 8 * 21) if version >= 2 and len(raw) >= 64 + 8 * 24 else (0, 0, 0)     # zero on pages of a library without send / recv
            # appended in round 2 (same version: older libraries leave these words zero): host-path calls / bytes / zero-copy / pipelined,
            # copy-engine launches, generic-reduction launches
            ext = struct.unpack_from("<6Q", raw, 64 + 8 * 24) if version >= 2 and len(raw) >= 64 + 8 * 30 else (0,) * 6
            pad = (0,) * (len(COLL_OPS) - nops)
            pages.append({"pid": pid, "rank": rank, "nranks": nranks, "device": device, "nvls": nvls, "version": version,
                          "calls": vals[0:nops] + pad, "bytes": vals[nops:2 * nops] + pad,
                          "algo_calls": vals[2 * nops:2 * nops + 7], "kernel_launches": vals[2 * nops + 7], "staged_calls": vals[2 * nops + 8],
                          "p2p_sends": p2p[0], "p2p_recvs": p2p[1], "p2p_bytes": p2p[2],
                          "host_calls": ext[0], "host_bytes": ext[1], "host_zero_copy": ext[2], "host_pipelined": ext[3],
                          "bulk_launches": ext[4], "generic_launches": ext[5]})
        except OSError:
            continue
    return pages


class MetricServer:
    def __init__(self, nvml, collection_interval_ms: int = 30000, port: int = 2112, pod_resources_socket: str = POD_RESOURCES_SOCKET,
                 registry: Optional[CollectorRegistry] = None, coll_stats_glob: str = "/dev/shm/b200coll.*", now=time.time):
        self.nvml, self.interval_ms, self.port, self.socket = nvml, collection_interval_ms, port, pod_resources_socket
        self.registry = registry or CollectorRegistry()
        self.coll_stats_glob = coll_stats_glob
        self._now = now
        self.last_reset = now()
        self.gpu_devices: dict = {}        # "nvidiaN" -> DeviceInfo
        self._stop = threading.Event()
        g = lambda name, help_, labels: Gauge(name, help_, labels, registry=self.registry)
        self.duty_cycle_node = g("duty_cycle_gpu_node", "Percent of time when the GPU was actively processing", _NODE_LABELS)
        self.memory_total_node = g("memory_total_gpu_node", "Total memory available on the GPU in bytes", _NODE_LABELS)
        self.memory_used_node = g("memory_used_gpu_node", "Allocated GPU memory in bytes", _NODE_LABELS)
        self.duty_cycle = g("duty_cycle", "Percent of time when the GPU was actively processing", _CONTAINER_LABELS)
        self.memory_total = g("memory_total", "Total memory available on the GPU in bytes", _CONTAINER_LABELS)
        self.memory_used = g("memory_used", "Allocated GPU memory in bytes", _CONTAINER_LABELS)
        self.requests = g("request", "Number of accelerator devices requested by the container", ["namespace", "pod", "container", "resource_name"])
        self.coll_calls = g("b200coll_calls", "Collective calls issued through libb200coll", ["pid", "rank", "op"])
        self.coll_bytes = g("b200coll_bytes", "Bytes moved by libb200coll collectives", ["pid", "rank", "op"])
        self.coll_algo = g("b200coll_algo_calls", "libb200coll calls per chosen algorithm", ["pid", "rank", "algo"])
        self.coll_p2p_calls = g("b200coll_p2p_calls", "Point-to-point operations issued through libb200coll", ["pid", "rank", "dir"])
        self.coll_p2p_bytes = g("b200coll_p2p_bytes", "Bytes sent plus received by libb200coll point-to-point operations", ["pid", "rank"])
        self.coll_host_calls = g("b200coll_host_calls", "End-to-end host all-reduces (b200collAllReduceHost) by path", ["pid", "rank", "path"])
        self.coll_host_bytes = g("b200coll_host_bytes", "Input bytes of the end-to-end host all-reduces", ["pid", "rank"])
        self.coll_family = g("b200coll_kernel_family_launches", "Launches of the copy-engine (bulk) and generic-reduction kernels", ["pid", "rank", "family"])
        self._all = [self.duty_cycle_node, self.memory_total_node, self.memory_used_node, self.duty_cycle, self.memory_total, self.memory_used,
                     self.requests, self.coll_calls, self.coll_bytes, self.coll_algo, self.coll_p2p_calls, self.coll_p2p_bytes,
                     self.coll_host_calls, self.coll_host_bytes, self.coll_family]

    def discover_gpu_devices(self) -> None:
        self.gpu_devices = {}
        for i in range(self.nvml.device_count()):
            info = self.nvml.device(i)
            self.gpu_devices[f"nvidia{info.minor}"] = info
            log.info("Found device nvidia%d for metrics collection", info.minor)

    def average_gpu_utilization(self, uuid: str, window_s: int = DUTY_CYCLE_WINDOW_S) -> int:
        since_us = int((self._now() - window_s) * 1e6)
        util = self.nvml.average_usage(uuid, since_us)
        if util > 100:
            raise ValueError(f"utilization ({util}) is out of range")
        return util

    def _metrics_info(self, device: str):
        info = self.gpu_devices.get(device)
        if info is None:
            raise KeyError(f"device {device} not found")
        fresh = self.nvml.device(info.index)
        return self.average_gpu_utilization(fresh.uuid), fresh

    def reset_if_needed(self) -> bool:
        if self._now() > self.last_reset + RESET_INTERVAL_S:
            for gauge in self._all:
                gauge.clear()
            self.last_reset = self._now()
            return True
        return False

    def update_metrics(self, container_devices: dict) -> None:
        self.reset_if_needed()
        for (ns, pod, ctr), devices in container_devices.items():
            self.requests.labels(ns, pod, ctr, GPU_RESOURCE_NAME).set(len(devices))
            for dev in devices:
                try:
                    duty, info = self._metrics_info(dev)
                except Exception as e:
                    log.info("Error calculating duty cycle for device: %s: %s. Skipping this device", dev, e)
                    continue
                labels = (ns, pod, ctr, "nvidia", info.uuid, info.name)
                self.duty_cycle.labels(*labels).set(duty)
                self.memory_total.labels(*labels).set(info.mem_total)
                self.memory_used.labels(*labels).set(info.mem_used)
        for dev in self.gpu_devices:
            try:
                duty, info = self._metrics_info(dev)
            except Exception as e:
                log.info("Error calculating duty cycle for device: %s: %s. Skipping this device", dev, e)
                continue
            labels = ("nvidia", info.uuid, info.name)
            self.duty_cycle_node.labels(*labels).set(duty)
            self.memory_total_node.labels(*labels).set(info.mem_total)
            self.memory_used_node.labels(*labels).set(info.mem_used)
        for page in read_coll_stats_pages(self.coll_stats_glob):
            pid, rank = str(page["pid"]), str(page["rank"])
            for i, op in enumerate(COLL_OPS):
                self.coll_calls.labels(pid, rank, op).set(page["calls"][i])
                self.coll_bytes.labels(pid, rank, op).set(page["bytes"][i])
            for i, algo in enumerate(COLL_ALGOS):
                self.coll_algo.labels(pid, rank, algo).set(page["algo_calls"][i])
            self.coll_p2p_calls.labels(pid, rank, "send").set(page["p2p_sends"])
            self.coll_p2p_calls.labels(pid, rank, "recv").set(page["p2p_recvs"])
            self.coll_p2p_bytes.labels(pid, rank).set(page["p2p_bytes"])
            for path, key in (("total", "host_calls"), ("zero_copy", "host_zero_copy"), ("pipelined", "host_pipelined")):
                self.coll_host_calls.labels(pid, rank, path).set(page[key])
            self.coll_host_bytes.labels(pid, rank).set(page["host_bytes"])
            self.coll_family.labels(pid, rank, "bulk").set(page["bulk_launches"])
            self.coll_family.labels(pid, rank, "generic").set(page["generic_launches"])

    def collect_once(self) -> None:
        try:
            devices = get_devices_for_all_containers(self.socket)
        except Exception as e:
            log.error("Failed to get devices for containers: %s", e)
            return
        self.update_metrics(devices)

    def start(self, serve_http: bool = True) -> None:
        log.info("Starting metrics server")
        log.info("nvml initialized successfully. Driver version: %s", self.nvml.driver_version())
        self.discover_gpu_devices()
        if serve_http:
            start_http_server(self.port, registry=self.registry)
        threading.Thread(target=self._loop, daemon=True).start()

    def _loop(self) -> None:
        while not self._stop.wait(self.interval_ms / 1000.0):
            self.collect_once()

    def stop(self) -> None:
        self._stop.set()
