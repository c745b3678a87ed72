"""Small helpers (reference: pkg/gpu/nvidia/util/util.go:29-70)."""
from __future__ import annotations

import os
import re

_DEV_PATH_RE = re.compile(r"^/dev/(nvidia[0-9]+)$")


def device_name_from_path(path: str) -> str:
    m = _DEV_PATH_RE.match(path)
    if not m:
        raise ValueError(f"path ({path}) is not a valid GPU device path")
    return m.group(1)


def node_name() -> str:
    """NODE_NAME from the downward API (the reference asks the GCE metadata server; SURVEY A.4 says don't)."""
    return os.environ.get("NODE_NAME") or os.uname().nodename
