"""Allocate hook for the collective-transport payload.

The reference's transport installers drop libnccl-net.so & friends into the host dir that Allocate already
mounts at /usr/local/nvidia (reference: cmd/nvidia_gpu/nvidia_gpu.go:46-49,81-83; consumers set
LD_LIBRARY_PATH=/usr/local/nvidia/lib64, gpudirect-rdma/nccl-test-a4.yaml:42-68) and ship an env-profile script
(`nccl-env-profile.sh`, gpudirect-tcpxo/README.md:25-35). Here the plugin exports the profile itself, so a pod needs
no wrapper script: B200COLL_* knobs plus an LD_LIBRARY_PATH hint, gated by GPUConfig.Transport.
"""
from __future__ import annotations

DEFAULT_ENV = {
    "B200COLL_LIB_DIR": "/usr/local/nvidia/lib64",
    "B200COLL_LIB": "/usr/local/nvidia/lib64/libb200coll.so",
    "B200COLL_NVLS": "-1",          # probe; 0 forces P2P paths (NCCL_NVLS_ENABLE analogue)
    "B200COLL_ALGO": "auto",        # NCCL_ALGO analogue: auto|ll|oneshot|twoshot|nvls
    "B200COLL_TIMEOUT_MS": "600000",
    "B200COLL_DEBUG": "WARN",
}


def env_profile(cfg) -> dict:
    env = dict(DEFAULT_ENV)
    env["B200COLL_LIB_DIR"] = cfg.lib_dir_container
    env["B200COLL_LIB"] = cfg.lib_dir_container.rstrip("/") + "/libb200coll.so"
    env["LD_LIBRARY_PATH"] = cfg.lib_dir_container
    env.update({str(k): str(v) for k, v in (cfg.env or {}).items()})
    return env


def apply(cfg, mount_paths: list, resp) -> None:
    """Add the transport env (and, if the lib dir is not under an existing mount, a read-only mount) to one
    ContainerAllocateResponse."""
    if not cfg or cfg.name != "b200coll":
        return
    for k, v in env_profile(cfg).items():
        if k not in resp.envs:
            resp.envs[k] = v
    covered = any(cfg.lib_dir_host == m.host_path or cfg.lib_dir_host.startswith(m.host_path.rstrip("/") + "/") for m in mount_paths)
    if not covered:
        resp.mounts.add(host_path=cfg.lib_dir_host, container_path=cfg.lib_dir_container, read_only=True)
