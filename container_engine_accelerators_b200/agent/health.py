"""GPU health checker: NVML Xid critical events -> Unhealthy devices (to ListAndWatch), a k8s Event, and the Node
condition `XidCriticalError` with heartbeat and bootID-based clearing.

Contract: reference pkg/gpu/nvidia/health_check/health_checker.go:65-468:
  * always-critical Xid 48 + XID_CONFIG extras mark devices Unhealthy; monitor-only set {48,63,64,79,119,120,123,140}
    only maintains the Node condition
  * Event: Warning / XIDError / "Caught XID error, XID=%d" / component nvidia-gpu-device-plugin
  * Condition: Type XidCriticalError, Status "True", Reason = JSON object {"<xid>":true,...} sorted keys,
    Message = bootID; a repeated Xid writes nothing
  * start-up: drop the condition iff stored bootID != current bootID (both non-empty); retry 1,2,4..30 s for <= 2 min
  * event without a UUID => every device Unhealthy; "Not Supported" at registration => GPU treated as always healthy
Differences: node name comes from NODE_NAME (downward API), not GCE metadata (A.4); the health hand-off never blocks.
"""
from __future__ import annotations

import json
import logging
import threading
import time
from typing import Optional

from . import kube as kubemod
from . import nvml as nvmlmod
from . import protos
from .mig import Device

log = logging.getLogger("b200-device-plugin")

XID_CONDITION_TYPE = "XidCriticalError"
EVENT_SOURCE = "nvidia-gpu-device-plugin"
MONITOR_XIDS = (48, 63, 64, 79, 119, 120, 123, 140)
ALWAYS_CRITICAL = (48,)
RESET_TIMEOUT_S = 120.0
HEARTBEAT_S = 60.0
WAIT_MS = 5000


class GPUHealthChecker:
    def __init__(self, devices: dict, report, critical_xids: list, kube: Optional[kubemod.KubeClient], nvml, node_name: str,
                 sleep=time.sleep, wait_ms: int = WAIT_MS):
        self.devices = {k: Device(v.id, v.health, v.numa_node) for k, v in devices.items()}   # private copy
        self.report = report                      # callable(Device) -> bool, never blocks
        self.kube, self.nvml, self.node_name = kube, nvml, node_name
        self.health_critical = set(int(c) for c in critical_xids) | set(ALWAYS_CRITICAL)
        self.monitor = set(MONITOR_XIDS)
        self.uuid_of: dict = {}                   # device id -> (gpu uuid, gi, ci)
        self._stop = threading.Event()
        self._threads: list = []
        self._event_set = None
        self._sleep = sleep
        self.wait_ms = wait_ms

    # ------------------------------------------------------------------ node condition
    def _update_node_status(self, mutate, attempts: int = 4) -> bool:
        """GET the Node, let `mutate(node)` edit it (return False for "nothing to write"), PUT the status. The PUT carries the
        resourceVersion of the GET, so a write that raced with somebody else's (the kubelet's status updates) is refused with 409:
        re-read and try again instead of losing the condition (the reference logs the error and gives up, health_checker.go:338-343)."""
        for attempt in range(attempts):
            node = self.kube.get_node(self.node_name)
            if mutate(node) is False:
                return False
            try:
                self.kube.update_node_status(node)
                return True
            except kubemod.KubeError as e:
                if e.status != 409 or attempt == attempts - 1:
                    raise
                log.info("Node %s changed under us (conflict); retrying the status update", self.node_name)
        return False

    def reset_xid_condition(self) -> bool:
        """Returns True if a condition was removed."""
        def mutate(node):
            status = node.setdefault("status", {})
            boot_id = (status.get("nodeInfo") or {}).get("bootID", "")
            conds = status.get("conditions") or []
            kept = []
            for c in conds:
                if c.get("type") == XID_CONDITION_TYPE and c.get("status") == "True":
                    last = c.get("message", "")
                    if boot_id and last and boot_id != last:
                        continue        # rebooted since the fault: auto-repair happened
                kept.append(c)
            if len(kept) == len(conds):
                return False
            status["conditions"] = kept
        if self._update_node_status(mutate):
            log.info("Successfully removed XIDCriticalError condition from node %s.", self.node_name)
            return True
        log.info("XIDCriticalError condition doesn't exist for node %s.", self.node_name)
        return False

    def reset_xid_condition_with_backoff(self, timeout_s: float = RESET_TIMEOUT_S) -> bool:
        backoff, deadline = 1.0, time.monotonic() + timeout_s
        while not self._stop.is_set():
            try:
                self.reset_xid_condition()
                return True
            except Exception as e:
                if time.monotonic() + backoff > deadline:
                    log.error("Timeout resetting XID condition after %.0f s.", timeout_s)
                    return False
                log.error("Failed to reset XID condition, will retry in %.0fs. Error: %s", backoff, e)
                self._sleep(backoff)
                backoff = min(backoff * 2, 30.0)
        return False

    def monitor_xid_event(self, xid: int) -> None:
        if xid not in self.monitor or self.kube is None:
            return

        def mutate(node):
            status = node.setdefault("status", {})
            conds = status.setdefault("conditions", [])
            for c in conds:
                if c.get("type") == XID_CONDITION_TYPE:
                    try:
                        reason = json.loads(c.get("reason") or "{}")
                    except ValueError:
                        log.error("Can't decode the value of condition.Reason %s", c.get("reason"))
                        return False
                    if str(xid) in reason:
                        log.info("XIDCriticalError condition already includes this XID %d, skip", xid)
                        return False
                    reason[str(xid)] = True
                    c["reason"] = json.dumps(reason, sort_keys=True, separators=(",", ":"))
                    return True
            now = kubemod.now_rfc3339()
            conds.append({"type": XID_CONDITION_TYPE, "status": "True", "lastHeartbeatTime": now, "lastTransitionTime": now,
                          "reason": json.dumps({str(xid): True}, separators=(",", ":")), "message": (status.get("nodeInfo") or {}).get("bootID", "")})
            return True
        try:
            if self._update_node_status(mutate):
                log.info("Successfully add XIDCriticalError condition on node %s.", self.node_name)
        except Exception as e:
            log.error("Failed to update node %s status to add XIDCriticalError condition: %s", self.node_name, e)

    def update_last_heartbeat(self) -> bool:
        if self.kube is None:
            return False

        def mutate(node):
            modified = False
            for c in (node.get("status") or {}).get("conditions") or []:
                if c.get("type") == XID_CONDITION_TYPE and c.get("status") == "True":
                    c["lastHeartbeatTime"] = kubemod.now_rfc3339()
                    modified = True
            return modified
        try:
            return self._update_node_status(mutate)
        except Exception as e:
            log.error("Failed to update node %s status to update XIDCondition heartbeat: %s", self.node_name, e)
            return False

    def record_xid_event(self, xid: int) -> None:
        if self.kube is None:
            return
        node = self.kube.get_node(self.node_name)
        involved = {"kind": "Node", "name": self.node_name, "uid": (node.get("metadata") or {}).get("uid", ""), "apiVersion": "v1"}
        self.kube.create_event("default", involved, "Warning", "XIDError", f"Caught XID error, XID={xid}", EVENT_SOURCE)

    # ------------------------------------------------------------------ event handling
    def catch_error(self, ev: nvmlmod.XidEvent) -> None:
        if ev.event_type != 8:     # nvmlEventTypeXidCriticalError
            log.info("Skip error Xid=%d as it is not Xid Critical", ev.xid)
            return
        try:
            self.record_xid_event(ev.xid)
        except Exception as e:
            log.error("Failed to record XID=%d for node %s with err %s", ev.xid, self.node_name, e)
        self.monitor_xid_event(ev.xid)
        if ev.xid not in self.health_critical:
            log.info("Health checker is skipping Xid %d error", ev.xid)
            return
        if not ev.uuid:
            log.error("XidCriticalError: Xid=%d, All devices will go unhealthy.", ev.xid)
            for d in self.devices.values():
                d.health = protos.UNHEALTHY
                self.report(Device(d.id, d.health, d.numa_node))
            return
        found = False
        for d in self.devices.values():
            ident = self.uuid_of.get(d.id)
            if ident is None:
                continue
            gpu, gi, ci = ident
            if gpu == ev.uuid and gi == ev.gpu_instance_id and (ci == ev.compute_instance_id or gi != nvmlmod.NOT_MIG):
                log.error("XidCriticalError: Xid=%d on Device=%s, uuid=%s, the device will go unhealthy.", ev.xid, d.id, gpu)
                d.health = protos.UNHEALTHY
                self.report(Device(d.id, d.health, d.numa_node))
                found = True
        if not found:
            log.error("XidCriticalError: Xid=%d on unknown device.", ev.xid)

    # ------------------------------------------------------------------ lifecycle
    def _index_devices(self) -> list:
        """Map plugin device ids to (gpu uuid, gi, ci); returns NVML indices to register."""
        indices = []
        for i in range(self.nvml.device_count()):
            info = self.nvml.device(i)
            name = f"nvidia{info.minor}"
            matched = False
            if name in self.devices:
                self.uuid_of[name] = (info.uuid, nvmlmod.NOT_MIG, nvmlmod.NOT_MIG)
                matched = True
            prefix = name + "/gi"
            for dev_id in self.devices:
                if dev_id.startswith(prefix):
                    try:
                        gi = int(dev_id[len(prefix):].split("/")[0])
                    except ValueError:
                        continue
                    self.uuid_of[dev_id] = (info.uuid, gi, 0)
                    matched = True
            if matched:
                indices.append(i)
            else:
                log.warning("Ignoring device %s for health check.", name)
        return indices

    def start(self, background: bool = True) -> None:
        if self.kube is not None and background:
            t = threading.Thread(target=self.reset_xid_condition_with_backoff, daemon=True); t.start(); self._threads.append(t)
            t = threading.Thread(target=self._heartbeat_loop, daemon=True); t.start(); self._threads.append(t)
        log.info("Starting GPU Health Checker")
        indices = self._index_devices()
        self._event_set = self.nvml.events_open()
        for i in indices:
            if not self.nvml.events_register(self._event_set, i):
                log.warning("Warning: GPU index %d is too old to support healthchecking. It will always be marked healthy.", i)
        if background:
            t = threading.Thread(target=self.listen_to_events, daemon=True); t.start(); self._threads.append(t)

    def _heartbeat_loop(self) -> None:
        while not self._stop.is_set():
            self.update_last_heartbeat()
            self._stop.wait(HEARTBEAT_S)

    def poll_once(self) -> bool:
        ev = self.nvml.events_wait(self._event_set, self.wait_ms)
        if ev is None:
            return False
        self.catch_error(ev)
        return True

    def listen_to_events(self) -> None:
        while not self._stop.is_set():
            try:
                self.poll_once()
            except Exception as e:
                log.error("GPUHealthChecker listen error: %s", e)
                self._stop.wait(1.0)

    def stop(self) -> None:
        self._stop.set()
        for t in self._threads:            # the listener may be inside events_wait(): let it return before freeing the set
            if t is not threading.current_thread():
                t.join(timeout=self.wait_ms / 1000.0 + 1.0)
        if self._event_set is not None:
            try:
                self.nvml.events_close(self._event_set)
            except Exception:
                pass
            self._event_set = None
