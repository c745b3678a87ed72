"""The NVML seam: one interface, a native implementation (ctypes over build/agent/libb200agent_nvml.so) and an
in-package mock that "discovers" GPUs by listing a fake /dev (the reference's test technique,
pkg/gpu/nvidia/nvmlutil/nvml_mock.go:28-70; seam: nvmlutil.go:30-37). NUMA topology comes from sysfs
(<pciRoot>/<busid>/numa_node) exactly as the reference derives it (nvmlutil.go:88-151).
"""
from __future__ import annotations

import ctypes as C
import os
import re
from dataclasses import dataclass
from pathlib import Path
from typing import Optional, Protocol

NVIDIA_DEVICE_RE = re.compile(r"^nvidia[0-9]*$")   # note the '*': the reference's pattern (manager.go:55)
PCI_DEVICES_ROOT = "/sys/bus/pci/devices"
EVENT_TIMEOUT, NOT_SUPPORTED, NO_LIB, NO_SAMPLES = -2, -3, -1, -5
NOT_MIG = 0xFFFFFFFF


@dataclass
class DeviceInfo:
    index: int
    minor: int
    uuid: str = ""
    name: str = ""
    bus_id: str = ""
    mem_total: int = 0
    mem_used: int = 0
    mig_mode: int = -1


@dataclass
class XidEvent:
    uuid: str            # "" => no device attached to the event
    xid: int
    gpu_instance_id: int = NOT_MIG
    compute_instance_id: int = NOT_MIG
    event_type: int = 8  # nvmlEventTypeXidCriticalError


class NvmlError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


class NvmlOperations(Protocol):
    def init(self) -> None: ...
    def device_count(self) -> int: ...
    def device(self, index: int) -> DeviceInfo: ...
    def driver_version(self) -> str: ...
    def average_usage(self, uuid: str, since_us: int) -> int: ...


def numa_topology(bus_id: str, pci_root: str = PCI_DEVICES_ROOT) -> Optional[int]:
    """NUMA node of a GPU, or None when the platform reports none (< 0). Raises on unreadable sysfs."""
    bus = bus_id.split("\x00")[0]
    domain = bus.split(":", 1)[0]
    if len(domain) == 8 and domain.startswith("0000"):
        bus = bus[4:]           # NVML pads the domain to 8 hex digits; sysfs uses 4 (the reference strips blindly, nvmlutil.go:131)
    bus = bus.lower()
    path = os.path.join(pci_root, bus, "numa_node")
    try:
        text = Path(path).read_text()
    except OSError as e:
        raise NvmlError(0, f"failed to read NUMA information from {path!r} file: {e}") from e
    try:
        node = int(text.strip())
    except ValueError as e:
        raise NvmlError(0, f"error parsing value for NUMA node: {e}") from e
    return node if node >= 0 else None


# --------------------------------------------------------------------------------------------- native
class _DevInfoC(C.Structure):
    _fields_ = [("index", C.c_int), ("minor_number", C.c_int), ("uuid", C.c_char * 96), ("name", C.c_char * 96), ("bus_id", C.c_char * 32),
                ("mem_total", C.c_ulonglong), ("mem_used", C.c_ulonglong), ("mig_mode_current", C.c_int), ("mig_mode_pending", C.c_int)]


class _EventC(C.Structure):
    _fields_ = [("uuid", C.c_char * 96), ("event_type", C.c_ulonglong), ("event_data", C.c_ulonglong), ("gpu_instance_id", C.c_uint),
                ("compute_instance_id", C.c_uint)]


def native_lib_path() -> Optional[str]:
    root = Path(__file__).resolve().parents[2]
    for cand in (os.environ.get("B200AGENT_NATIVE_LIB", ""), str(root / "build" / "agent" / "libb200agent_nvml.so"), "/usr/local/lib/libb200agent_nvml.so"):
        if cand and os.path.exists(cand):
            return cand
    return None


class NativeNvml:
    """ctypes front for libb200agent_nvml.so (which dlopens libnvidia-ml.so.1, or $B200AGENT_NVML_LIB)."""

    def __init__(self, lib_path: Optional[str] = None):
        path = lib_path or native_lib_path()
        if not path:
            raise NvmlError(NO_LIB, "libb200agent_nvml.so not built (make -C agent/native)")
        self.L = L = C.CDLL(path)
        L.b200nvml_last_error.restype = C.c_char_p
        L.b200nvml_device_info_get.argtypes = [C.c_int, C.POINTER(_DevInfoC)]
        L.b200nvml_average_usage.argtypes = [C.c_char_p, C.c_ulonglong, C.POINTER(C.c_uint)]
        L.b200nvml_driver_version.argtypes = [C.c_char_p, C.c_uint]
        L.b200nvml_events_open.argtypes = [C.POINTER(C.c_void_p)]
        L.b200nvml_events_register_xid.argtypes = [C.c_void_p, C.c_int]
        L.b200nvml_events_wait.argtypes = [C.c_void_p, C.c_uint, C.POINTER(_EventC)]
        L.b200nvml_events_close.argtypes = [C.c_void_p]
        L.b200nvml_event_type_xid.restype = C.c_ulonglong

    def _ck(self, rc: int, what: str) -> None:
        if rc != 0:
            raise NvmlError(rc, f"{what}: {self.L.b200nvml_last_error().decode(errors='replace')}")

    def init(self) -> None:
        self._ck(self.L.b200nvml_init(), "nvml init")

    def shutdown(self) -> None:
        self.L.b200nvml_shutdown()

    def device_count(self) -> int:
        n = C.c_int()
        self._ck(self.L.b200nvml_device_count(C.byref(n)), "failed to get devices count")
        return n.value

    def device(self, index: int) -> DeviceInfo:
        d = _DevInfoC()
        self._ck(self.L.b200nvml_device_info_get(index, C.byref(d)), f"failed to get the device handle for index {index}")
        return DeviceInfo(index, d.minor_number, d.uuid.decode(), d.name.decode(), d.bus_id.decode(), d.mem_total, d.mem_used, d.mig_mode_current)

    def driver_version(self) -> str:
        buf = C.create_string_buffer(96)
        self._ck(self.L.b200nvml_driver_version(buf, len(buf)), "driver version")
        return buf.value.decode()

    def average_usage(self, uuid: str, since_us: int) -> int:
        u = C.c_uint()
        self._ck(self.L.b200nvml_average_usage(uuid.encode(), since_us, C.byref(u)), "average usage")
        return u.value

    # ---- Xid events
    def events_open(self):
        h = C.c_void_p()
        self._ck(self.L.b200nvml_events_open(C.byref(h)), "event set create")
        return h

    def events_register(self, handle, index: int) -> bool:
        """True if registered; False when the GPU does not support Xid events ("Not Supported" => always healthy)."""
        rc = self.L.b200nvml_events_register_xid(handle, index)
        if rc == NOT_SUPPORTED:
            return False
        self._ck(rc, "register events")
        return True

    def events_wait(self, handle, timeout_ms: int) -> Optional[XidEvent]:
        ev = _EventC()
        rc = self.L.b200nvml_events_wait(handle, timeout_ms, C.byref(ev))
        if rc == EVENT_TIMEOUT:
            return None
        self._ck(rc, "event wait")
        return XidEvent(ev.uuid.decode(), int(ev.event_data), ev.gpu_instance_id, ev.compute_instance_id, int(ev.event_type))

    def events_close(self, handle) -> None:
        self.L.b200nvml_events_close(handle)


# --------------------------------------------------------------------------------------------- mock
class MockNvml:
    """Counts nvidiaN files in a temp /dev; minor == index; bus id settable per test."""

    def __init__(self, dev_dir: str, bus_id: str = "", mem_total: int = 80 << 30, uuids: Optional[list] = None, name: str = "NVIDIA B200",
                 driver: str = "580.159.03", bus_ids: Optional[list] = None):
        self.dev_dir, self.bus_id, self.mem_total, self.uuids, self.name, self.driver = dev_dir, bus_id, mem_total, uuids, name, driver
        self.bus_ids = bus_ids                      # per-index bus ids (two-socket hosts); falls back to `bus_id`
        self.utilisation: dict = {}
        self.events: list = []
        self.events_supported = True

    def init(self) -> None:
        pass

    def device_count(self) -> int:
        try:
            return sum(1 for e in os.scandir(self.dev_dir) if not e.is_dir() and NVIDIA_DEVICE_RE.match(e.name))
        except OSError:
            return 0

    def device(self, index: int) -> DeviceInfo:
        uuid = self.uuids[index] if self.uuids and index < len(self.uuids) else f"GPU-mock-{index}"
        bus = self.bus_ids[index] if self.bus_ids and index < len(self.bus_ids) else self.bus_id
        return DeviceInfo(index, index, uuid, self.name, bus, self.mem_total, 0, 0)

    def driver_version(self) -> str:
        return self.driver

    def average_usage(self, uuid: str, since_us: int) -> int:
        if uuid not in self.utilisation:
            raise NvmlError(NO_SAMPLES, "no samples")
        return self.utilisation[uuid]

    def events_open(self):
        return object()

    def events_register(self, handle, index: int) -> bool:
        return self.events_supported

    def events_wait(self, handle, timeout_ms: int) -> Optional[XidEvent]:
        return self.events.pop(0) if self.events else None

    def events_close(self, handle) -> None:
        pass
