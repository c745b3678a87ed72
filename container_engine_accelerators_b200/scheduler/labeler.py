"""Topology node labeler.

Source `gce-metadata` reproduces the reference (gke-topology-scheduler/label-nodes-daemon.py:26-69): instance name +
attributes/physical_host ("/cluster/rack/host") -> labels topology.gke.io/{cluster,rack,host}, every 600 s.
Source `nvml` is the single-box analogue: on one NVSwitch host every peer is equidistant, so the useful
labels are the GPU count/model, the NUMA spread and an NVLink-domain id shared by every GPU behind the same switch fabric.
"""
from __future__ import annotations

import argparse
import hashlib
import logging
import sys
import time

import requests

from ..agent import nvml as nvmlmod
from ..agent.kube import KubeClient
from ..agent.util import node_name

log = logging.getLogger("b200-topology-labeler")

METADATA_URL = "http://metadata.google.internal/computeMetadata/v1/instance"
HEADERS = {"Metadata-Flavor": "Google"}
PERIOD_S = 600


def labels_from_physical_host(physical_host: str) -> dict:
    parts = physical_host.split("/")[1:]
    if len(parts) != 3:
        raise ValueError(f"unexpected physical_host {physical_host!r}, want /cluster/rack/host")
    cluster, rack, host = parts
    return {"topology.gke.io/cluster": cluster, "topology.gke.io/rack": rack, "topology.gke.io/host": host}


def update_node_labels_from_metadata(kube: KubeClient, metadata_url: str = METADATA_URL, session=requests) -> dict:
    r = session.get(f"{metadata_url}/name", headers=HEADERS, timeout=10)
    if r.status_code != 200:
        log.error("Node name not found")
        return {}
    name = r.text
    r = session.get(f"{metadata_url}/attributes/physical_host", headers=HEADERS, timeout=10)
    if r.status_code != 200:
        log.error("physical host not found")
        return {}
    labels = labels_from_physical_host(r.text)
    kube.patch_node_labels(name, labels)
    log.info("Updated labels on node %s: %s", name, labels)
    return labels


def labels_from_nvml(api, pci_root: str = nvmlmod.PCI_DEVICES_ROOT) -> dict:
    count = api.device_count()
    infos = [api.device(i) for i in range(count)]
    numa = set()
    for info in infos:
        try:
            n = nvmlmod.numa_topology(info.bus_id, pci_root) if info.bus_id else None
        except nvmlmod.NvmlError:
            n = None
        if n is not None:
            numa.add(n)
    domain = hashlib.sha1(",".join(sorted(i.uuid for i in infos)).encode()).hexdigest()[:12] if infos else ""
    model = infos[0].name.replace(" ", "-") if infos else ""
    return {"b200.gke.io/gpu-count": str(count), "b200.gke.io/gpu-model": model, "b200.gke.io/numa-nodes": str(len(numa)),
            "b200.gke.io/nvlink-domain": domain}


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="b200-topology-labeler")
    ap.add_argument("--source", choices=["gce-metadata", "nvml"], default="gce-metadata")
    ap.add_argument("--period", type=int, default=PERIOD_S)
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--kube-url", default="", help="API server URL (default: B200_KUBE_URL, else in-cluster)")
    ap.add_argument("--metadata-url", default=METADATA_URL, help="GCE metadata endpoint (tests, proxies)")
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(levelname).1s %(name)s] %(message)s")
    kube = KubeClient.from_env(args.kube_url)
    api = None
    if args.source == "nvml":
        api = nvmlmod.NativeNvml(); api.init()
    while True:
        log.info("Starting node update")
        try:
            if args.source == "gce-metadata":
                update_node_labels_from_metadata(kube, args.metadata_url)
            else:
                kube.patch_node_labels(node_name(), labels_from_nvml(api))
        except Exception as e:
            log.error("label update failed: %s", e)
        if args.once:
            return 0
        time.sleep(args.period)


if __name__ == "__main__":
    sys.exit(main())
