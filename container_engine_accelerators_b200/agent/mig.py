"""MIG device manager: discovers GPU-instance / compute-instance capability files and turns each partition
into an allocatable device with three DeviceSpecs (parent GPU, GI cap, CI cap).

Contract: reference pkg/gpu/nvidia/mig/mig.go:111-254. The size table is shared with the C++
partitioner through agent/native/mig_profiles.inc. Unlike the reference, the maps are guarded by the
manager's lock (mig.go:257-259 is unguarded).
"""
from __future__ import annotations

import logging
import os
import re
from dataclasses import dataclass
from pathlib import Path
from typing import Callable, Optional

from . import nvml as nvmlmod

log = logging.getLogger("b200-device-plugin")

_PROFILE_RE = re.compile(r'MIG_PROFILE\("([^"]+)",\s*(\d+),\s*(\d+),\s*"([^"]*)"\)')
_GPU_RE = re.compile(r"gpu([0-9]+)")
_GI_RE = re.compile(r"gi([0-9]+)")
_MINOR_RE = re.compile(r"DeviceFileMinor: ([0-9]+)")


@dataclass(frozen=True)
class MigProfile:
    size: str
    profile_id: int
    max_count: int
    families: tuple


def _load_profiles() -> dict:
    inc = Path(__file__).resolve().parents[2] / "agent" / "native" / "mig_profiles.inc"
    out = {}
    for m in _PROFILE_RE.finditer(inc.read_text()):
        out[m.group(1)] = MigProfile(m.group(1), int(m.group(2)), int(m.group(3)), tuple(m.group(4).split(",")))
    return out


PROFILES: dict = _load_profiles()


class MigError(RuntimeError):
    pass


@dataclass
class Spec:
    host_path: str
    container_path: str
    permissions: str = "mrw"


@dataclass
class Device:
    id: str
    health: str
    numa_node: Optional[int] = None


class MigDeviceManager:
    def __init__(self, dev_dir: str, proc_dir: str, topology_of: Optional[Callable[[int], Optional[int]]] = None):
        self.dev_dir, self.proc_dir = dev_dir, proc_dir
        self.partition_specs: dict = {}
        self.partitions: dict = {}
        self._topology_of = topology_of or (lambda idx: None)

    def list_partitions(self) -> dict:
        return self.partitions

    def device_spec(self, device_id: str) -> list:
        if device_id not in self.partition_specs:
            raise MigError(f"invalid allocation request with non-existing GPU partition: {device_id}")
        return self.partition_specs[device_id]

    def set_device_health(self, name: str, health: str, numa_node: Optional[int] = None) -> None:
        self.partitions[name] = Device(name, health, numa_node)

    def _num_gpus(self) -> int:
        try:
            return sum(1 for e in os.scandir(self.dev_dir) if not e.is_dir() and nvmlmod.NVIDIA_DEVICE_RE.match(e.name))
        except OSError as e:
            raise MigError(f"failed to read devices on node: {e}") from e

    def start(self, partition_size: str) -> None:
        if not partition_size:
            return
        prof = PROFILES.get(partition_size)
        if prof is None:
            raise MigError(f"{partition_size} is not a valid GPU partition size")
        specs: dict = {}
        parts: dict = {}
        cap_dir = os.path.join(self.proc_dir, "driver/nvidia/capabilities")
        try:
            cap_entries = sorted(os.listdir(cap_dir))
        except OSError as e:
            raise MigError(f"failed to read capabilities directory ({cap_dir}): {e}") from e
        partitioned_gpus = 0
        for entry in cap_entries:
            m = _GPU_RE.search(entry)
            if not m:
                continue
            gpu_id = m.group(1)
            partitioned_gpus += 1
            gi_base = os.path.join(cap_dir, entry, "mig")
            try:
                gi_entries = sorted(os.listdir(gi_base))
            except OSError as e:
                raise MigError(f"failed to read GPU instance capabilities dir ({gi_base}): {e}") from e
            count = 0
            for gi in gi_entries:
                if not _GI_RE.search(gi):
                    continue
                count += 1
                instance_id = f"nvidia{gpu_id}/{gi}"
                gi_minor = self._minor(os.path.join(gi_base, gi, "access"), "GPU instance")
                ci_minor = self._minor(os.path.join(gi_base, gi, "ci0", "access"), "compute instance")   # only ci0 is considered
                gpu_dev = os.path.join(self.dev_dir, f"nvidia{gpu_id}")
                gi_dev = os.path.join(self.dev_dir, "nvidia-caps", f"nvidia-cap{gi_minor}")
                ci_dev = os.path.join(self.dev_dir, "nvidia-caps", f"nvidia-cap{ci_minor}")
                for what, p in (("GPU device", gpu_dev), ("GPU instance device", gi_dev), ("Compute instance device", ci_dev)):
                    if not os.path.exists(p):
                        raise MigError(f"{what} ({p}) not found")
                log.info("Discovered GPU partition: %s", instance_id)
                specs[instance_id] = [Spec(gpu_dev, gpu_dev), Spec(gi_dev, gi_dev), Spec(ci_dev, ci_dev)]
                numa = None
                try:
                    numa = self._topology_of(int(gpu_id))
                except Exception as e:   # topology is advisory
                    log.error("unable to get topology for device with index %s: %s", gpu_id, e)
                parts[instance_id] = Device(instance_id, "Healthy", numa)
            if count != prof.max_count:
                raise MigError(f"Number of partitions ({count}) for GPU {gpu_id} does not match expected partition count ({prof.max_count})")
        num_gpus = self._num_gpus()
        if partitioned_gpus != num_gpus:
            raise MigError(f"Not all GPUs are partitioned as expected. Total number of GPUs: {num_gpus}, number of partitioned GPUs: {partitioned_gpus}")
        self.partition_specs, self.partitions = specs, parts

    @staticmethod
    def _minor(path: str, what: str) -> int:
        try:
            text = Path(path).read_text()
        except OSError as e:
            raise MigError(f"failed to read {what} access file ({path}): {e}") from e
        m = _MINOR_RE.search(text)
        if not m:
            raise MigError(f"unexpected contents in {what} access file({path})")
        return int(m.group(1))
