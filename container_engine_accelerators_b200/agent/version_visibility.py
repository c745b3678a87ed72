"""Publish the driver version as Node annotations via server-side apply.

Contract: reference pkg/gpu/nvidia/version_visibility/version_visibility.go:29-86 — keys
cloud.google.com/cuda.driver-version.{major,minor,revision,full}, FieldManager gpu-device-plugin, Force, other
annotations preserved (version_visibility_test.go:125-127).
"""
from __future__ import annotations

import logging

log = logging.getLogger("b200-device-plugin")

PREFIX = "cloud.google.com/cuda.driver-version."
MAJOR, MINOR, REVISION, FULL = PREFIX + "major", PREFIX + "minor", PREFIX + "revision", PREFIX + "full"
FIELD_MANAGER = "gpu-device-plugin"


def parse_driver_annotations(version: str) -> dict:
    parts = version.strip().split(".")
    if len(parts) < 2 or len(parts) > 3 or not all(p.isdigit() for p in parts):
        raise ValueError(f"unexpected driver version format: {version!r}")
    return {MAJOR: parts[0], MINOR: parts[1], REVISION: parts[2] if len(parts) == 3 else "", FULL: version.strip()}


def publish_driver_version_annotations(kube, node_name: str, driver_version: str) -> dict:
    ann = parse_driver_annotations(driver_version)
    kube.apply_node_annotations(node_name, ann, FIELD_MANAGER, force=True)
    log.info("published driver version %s on node %s", driver_version, node_name)
    return ann
