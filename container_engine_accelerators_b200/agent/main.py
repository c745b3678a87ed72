"""b200-device-plugin entry point.

Boot sequence follows reference cmd/nvidia_gpu/nvidia_gpu.go:78-186: parse flags and
gpu_config.json -> wait for /dev/nvidiactl + /dev/nvidia-uvm (5 s poll) -> NVML init -> manager.start() retried
every 5 s -> optional metrics / health checker / driver-version annotations -> serve (blocks).
    python -m container_engine_accelerators_b200.agent.main --enable-health-monitoring --enable-container-gpu-metrics
"""
from __future__ import annotations

import argparse
import logging
import os
import signal
import sys
import threading
import time

from . import health, kube, metrics, nvml, util, version_visibility
from .config import parse_gpu_config
from .manager import GPUManager, Mount

log = logging.getLogger("b200-device-plugin")

KUBELET_ENDPOINT = "kubelet.sock"
PLUGIN_ENDPOINT_PREFIX = "nvidiaGPU"
DEV_DIR, PROC_DIR = "/dev", "/proc"


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="b200-device-plugin")
    p.add_argument("--host-path", default="/home/kubernetes/bin/nvidia", help="host dir with NVIDIA libraries, binaries and the collective transport")
    p.add_argument("--container-path", default="/usr/local/nvidia", help="where --host-path is mounted in containers")
    p.add_argument("--host-vulkan-icd-path", default="/home/kubernetes/bin/nvidia/vulkan/icd.d")
    p.add_argument("--container-vulkan-icd-path", default="/etc/vulkan/icd.d")
    p.add_argument("--plugin-directory", default="/device-plugin")
    p.add_argument("--enable-container-gpu-metrics", action="store_true")
    p.add_argument("--enable-health-monitoring", action="store_true")
    p.add_argument("--gpu-metrics-port", type=int, default=2112)
    p.add_argument("--gpu-metrics-collection-interval", type=int, default=30000, help="ms")
    p.add_argument("--gpu-config", default="/etc/nvidia/gpu_config.json")
    p.add_argument("--publish-driver-version", action="store_true")
    p.add_argument("--dev-directory", default=DEV_DIR)
    p.add_argument("--proc-directory", default=PROC_DIR)
    # seams the conformance harness uses to run the agent as a real process against fake /dev, /proc, /sys trees (same names as the native binary)
    p.add_argument("--pci-root", default=nvml.PCI_DEVICES_ROOT, help=argparse.SUPPRESS)
    p.add_argument("--plugin-endpoint", default="", help=argparse.SUPPRESS)
    p.add_argument("--mps-control-bin", default=None, help=argparse.SUPPRESS)
    p.add_argument("--pod-resources-socket", default=metrics.POD_RESOURCES_SOCKET, help=argparse.SUPPRESS)
    p.add_argument("--coll-stats-dir", default="/dev/shm", help=argparse.SUPPRESS)
    p.add_argument("--gpu-check-interval", type=float, default=None, help=argparse.SUPPRESS)
    p.add_argument("--socket-check-interval", type=float, default=None, help=argparse.SUPPRESS)
    p.add_argument("--kube-url", default="", help="API server URL (default: in-cluster); B200_KUBE_URL is honoured too")
    p.add_argument("--status-only", action="store_true",
                   help="do not serve the kubelet API (the native b200-device-plugin does); only publish Kubernetes-side status: Xid Events + Node condition, driver-version annotations")
    p.add_argument("--preferred-allocation-policy", choices=["none", "spread", "packed"], default="none",
                   help="answer the kubelet's GetPreferredAllocation (NUMA-aligned; spread or pack shared GPUs). none = the reference's behaviour: no plugin options")
    p.add_argument("-v", "--verbosity", type=int, default=0)
    for glog_flag in ("--logtostderr", "--alsologtostderr"):                # accepted for drop-in compatibility with the Go binary's manifests
        p.add_argument(glog_flag, nargs="?", const="true", default="true", help=argparse.SUPPRESS)
    return p


def go_style_argv(argv: list) -> list:
    """The reference's binary is Go: `-enable-health-monitoring`, `-gpu-config=/x` (single dash). Accept that spelling too, so the
    same args line drives either implementation; `-v 3` and other one-letter flags are left alone."""
    out = []
    for a in argv:
        out.append("-" + a if len(a) > 2 and a[0] == "-" and a[1] != "-" and a[1].isalpha() and (a[2].isalpha() or a[2] == "-") else a)
    return out


def main(argv=None) -> int:
    args = build_parser().parse_args(go_style_argv(list(sys.argv[1:] if argv is None else argv)))
    logging.basicConfig(stream=sys.stderr, level=logging.DEBUG if args.verbosity >= 3 else logging.INFO, format="%(asctime)s %(levelname).1s %(name)s] %(message)s")
    log.info("device-plugin started")
    mounts = [Mount(args.host_path, args.container_path, True), Mount(args.host_vulkan_icd_path, args.container_vulkan_icd_path, True)]
    cfg = parse_gpu_config(args.gpu_config)
    try:
        cfg.add_health_critical_xid()
    except ValueError as e:
        log.error("failed to add HealthCriticalXid: %s", e)
    log.info("Using gpu config: %s", cfg)
    api = nvml.NativeNvml()
    seams = {k: v for k, v in (("gpu_check_interval", args.gpu_check_interval), ("socket_check_interval", args.socket_check_interval), ("mps_control_bin", args.mps_control_bin)) if v is not None}
    ngm = GPUManager(args.dev_directory, args.proc_directory, mounts, cfg, nvml=api, pci_root=args.pci_root, preferred_allocation_policy=args.preferred_allocation_policy, **seams)
    while True:
        try:
            ngm.check_device_paths()
            break
        except OSError:
            log.debug("nvidiactl / nvidia-uvm not present yet; waiting for the driver installer")
            time.sleep(5)
    log.info("Initializing nvml")
    api.init()
    while True:
        try:
            ngm.start()
            break
        except Exception as e:
            log.error("failed to start GPU device manager: %s", e)
            time.sleep(5)
    if args.enable_container_gpu_metrics:
        if cfg.gpu_partition_size:
            log.info("metrics are disabled when MIG partitioning is on")
        else:
            log.info("Starting metrics server on port: %d, collection interval: %d", args.gpu_metrics_port, args.gpu_metrics_collection_interval)
            try:
                metrics.MetricServer(api, args.gpu_metrics_collection_interval, args.gpu_metrics_port, pod_resources_socket=args.pod_resources_socket,
                                     coll_stats_glob=os.path.join(args.coll_stats_dir, "b200coll.*")).start()
            except Exception as e:
                log.error("failed to start metric server: %s", e)
    kc = None
    if args.enable_health_monitoring or args.publish_driver_version:
        try:
            kc = kube.KubeClient.from_env(args.kube_url)
        except Exception as e:
            log.error("failed to build kube client: %s", e)
    if args.enable_health_monitoring:
        report = (lambda d: True) if args.status_only else ngm.report_unhealthy
        hc = health.GPUHealthChecker(ngm.list_physical_devices(), report, ngm.list_health_critical_xid(), kc, api, util.node_name())
        try:
            hc.start()
        except Exception as e:
            log.error("failed to start GPU Health Checker: %s", e)
    if args.publish_driver_version and kc is not None:
        def publish():
            try:
                version_visibility.publish_driver_version_annotations(kc, util.node_name(), api.driver_version())
            except Exception as e:
                log.error("failed to publish driver version annotations: %s", e)
        threading.Thread(target=publish, daemon=True).start()
    # SIGTERM / SIGINT (pod deletion, rolling update): stop serving, remove the plugin socket, exit 0 — the kubelet sees the endpoint
    # go away cleanly instead of a dangling socket file
    stopping = threading.Event()

    def on_signal(signum, frame):
        log.info("received signal %d, shutting down", signum)
        stopping.set()
        ngm.stop()
    signal.signal(signal.SIGTERM, on_signal)
    signal.signal(signal.SIGINT, on_signal)
    if args.status_only:
        log.info("status-only mode: the kubelet-facing API is served by the native plugin")
        stopping.wait()
        return 0
    ngm.serve(args.plugin_directory, KUBELET_ENDPOINT, args.plugin_endpoint or f"{PLUGIN_ENDPOINT_PREFIX}-{int(time.time())}.sock")
    return 0


if __name__ == "__main__":
    sys.exit(main())
