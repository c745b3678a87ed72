"""GPU sharing (time-sharing / MPS): request validation and virtual -> physical device-id mapping.

Contract: reference pkg/gpu/nvidia/gpusharing/gpusharing.go:25-77 (error strings are asserted by its tests,
gpusharing_test.go:47,53). Unlike the reference there is no mutable package global: the strategy is passed in.
"""
from __future__ import annotations

import re

UNDEFINED = ""
TIME_SHARING = "time-sharing"
MPS = "mps"

_VGPU_DEFAULT = re.compile(r"nvidia([0-9]+)/vgpu([0-9]+)$")
_VGPU_MIG = re.compile(r"nvidia([0-9]+)/gi([0-9]+)/vgpu([0-9]+)$")
_VGPU_SUFFIX = re.compile(r"/vgpu([0-9]+)$")

ERR_TIME_SHARING = "invalid request for sharing GPU (time-sharing), at most 1 nvidia.com/gpu can be requested on GPU nodes"
ERR_MPS = "invalid request for sharing GPU (MPS), at most 1 nvidia.com/gpu can be requested on multi-GPU nodes"


class SharingError(ValueError):
    pass


def is_virtual_device_id(device_id: str) -> bool:
    return bool(_VGPU_DEFAULT.search(device_id) or _VGPU_MIG.search(device_id))


def validate_request(device_ids: list[str], physical_device_count: int, strategy: str) -> None:
    """Only multi-device requests whose first id is virtual are constrained: time-sharing never allows them,
    MPS allows them on single-GPU nodes (each MIG partition counts as a physical device)."""
    if len(device_ids) > 1 and is_virtual_device_id(device_ids[0]):
        if strategy == TIME_SHARING:
            raise SharingError(ERR_TIME_SHARING)
        if strategy == MPS and physical_device_count > 1:
            raise SharingError(ERR_MPS)


def virtual_to_physical_device_id(virtual_id: str) -> str:
    if not is_virtual_device_id(virtual_id):
        raise SharingError(f"virtual device ID ({virtual_id}) is not valid")
    return _VGPU_SUFFIX.split(virtual_id)[0]
