"""Node GPU configuration (/etc/nvidia/gpu_config.json) and XID_CONFIG parsing.

Behavioural contract: reference pkg/gpu/nvidia/manager.go:72-137 and cmd/nvidia_gpu/nvidia_gpu.go:59-99
(SURVEY Appendix A.2). New field `Transport` gates the libb200coll mount/env hook in Allocate.
"""
from __future__ import annotations

import json
import logging
import os
from dataclasses import dataclass, field
from typing import Optional

from . import sharing

log = logging.getLogger("b200-device-plugin")


class ConfigError(ValueError):
    pass


@dataclass
class GPUSharingConfig:
    strategy: str = sharing.UNDEFINED
    max_shared_clients_per_gpu: int = 0


@dataclass
class TransportConfig:
    """Which collective-transport payload Allocate should expose to containers."""
    name: str = ""                       # "" (none) | "b200coll"
    lib_dir_host: str = "/home/kubernetes/bin/nvidia/lib64"
    lib_dir_container: str = "/usr/local/nvidia/lib64"
    env: dict = field(default_factory=dict)


@dataclass
class GPUConfig:
    gpu_partition_size: str = ""
    max_time_shared_clients_per_gpu: int = 0     # deprecated spelling, wins when > 0
    sharing: GPUSharingConfig = field(default_factory=GPUSharingConfig)
    health_critical_xid: list = field(default_factory=list)
    transport: TransportConfig = field(default_factory=TransportConfig)

    # ------------------------------------------------------------------ parsing
    @classmethod
    def from_json(cls, text: str) -> "GPUConfig":
        raw = json.loads(text)
        if not isinstance(raw, dict):
            raise ConfigError("gpu config must be a JSON object")
        cfg = cls()
        cfg.gpu_partition_size = str(raw.get("GPUPartitionSize", "") or "")
        cfg.max_time_shared_clients_per_gpu = int(raw.get("MaxTimeSharedClientsPerGPU", 0) or 0)
        sh = raw.get("GPUSharingConfig") or {}
        cfg.sharing = GPUSharingConfig(str(sh.get("GPUSharingStrategy", "") or ""), int(sh.get("MaxSharedClientsPerGPU", 0) or 0))
        cfg.health_critical_xid = [int(x) for x in (raw.get("HealthCriticalXid") or [])]
        tr = raw.get("Transport") or {}
        if isinstance(tr, str):
            tr = {"Name": tr}
        cfg.transport = TransportConfig(str(tr.get("Name", "") or ""), str(tr.get("LibDirHost", TransportConfig.lib_dir_host)),
                                        str(tr.get("LibDirContainer", TransportConfig.lib_dir_container)), dict(tr.get("Env") or {}))
        return cfg

    def add_defaults_and_validate(self) -> None:
        """reference: manager.go:90-115. Sets the process-wide sharing strategy as the reference does (gpusharing.go:31)."""
        if self.max_time_shared_clients_per_gpu > 0:
            if self.sharing.strategy or self.sharing.max_shared_clients_per_gpu > 0:
                log.info("Both MaxTimeSharedClientsPerGPU and GPUSharingConfig are set, use the value of MaxTimeSharedClientsPerGPU")
            self.sharing = GPUSharingConfig(sharing.TIME_SHARING, self.max_time_shared_clients_per_gpu)
        elif self.sharing.strategy in (sharing.TIME_SHARING, sharing.MPS):
            if self.sharing.max_shared_clients_per_gpu <= 0:
                raise ConfigError("MaxSharedClientsPerGPU should be > 0 for time-sharing or mps GPU sharing strategies")
        elif self.sharing.strategy == sharing.UNDEFINED:
            if self.sharing.max_shared_clients_per_gpu > 0:
                raise ConfigError("GPU sharing strategy needs to be specified when MaxSharedClientsPerGPU > 0")
        else:
            raise ConfigError(f"invalid GPU Sharing strategy: {self.sharing.strategy}, should be one of time-sharing or mps")
        if self.transport.name not in ("", "b200coll"):
            raise ConfigError(f"invalid Transport: {self.transport.name}, should be empty or b200coll")

    def add_health_critical_xid(self, env: Optional[dict] = None) -> None:
        """XID_CONFIG="a, b,c" (reference: manager.go:117-137). A bad token raises; the caller logs and continues."""
        xid_config = (env if env is not None else os.environ).get("XID_CONFIG", "")
        if not xid_config:
            log.info("There is no Xid config specified")
            return
        out = []
        for tok in xid_config.split(","):
            tok = tok.strip()
            try:
                out.append(int(tok))
            except ValueError as e:
                raise ConfigError(f"Invalid HealthCriticalXid input : {e}") from e
        self.health_critical_xid = out


def parse_gpu_config(path: str) -> GPUConfig:
    """Missing file => empty config. Parse/validation failure => log and fall back to the empty config
    (reference: cmd/nvidia_gpu/nvidia_gpu.go:59-76,89-94)."""
    cfg = GPUConfig()
    try:
        with open(path) as f:
            text = f.read()
    except FileNotFoundError:
        log.info("No GPU config file (%s); using defaults", path)
        cfg.add_defaults_and_validate()
        return cfg
    except OSError as e:
        log.error("unable to read gpu config file %s: %s", path, e)
        cfg.add_defaults_and_validate()
        return cfg
    try:
        parsed = GPUConfig.from_json(text)
        parsed.add_defaults_and_validate()
        return parsed
    except (ValueError, TypeError) as e:
        log.error("failed to parse GPU config file %s: %s; falling back to default GPU config", path, e)
        cfg.add_defaults_and_validate()
        return cfg
