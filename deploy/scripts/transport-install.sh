#!/bin/bash
# Transport-installer payload step: run the image's own installer, then copy its lib64 tree onto the host dir that the
# device plugin mounts into every GPU container. Behaviour: reference gpudirect-rdma/nccl-rdma-installer.yaml:70-77,
# gpudirect-tcpx/nccl-tcpx-installer.yaml:57-64, gpudirect-tcpxo/nccl-tcpxo-installer.yaml:83-91.
set -eu
SRC="${TRANSPORT_SRC_DIR:?set TRANSPORT_SRC_DIR (e.g. /var/lib/tcpxo/lib64)}"
DST="${NCCL_INSTALL_DIR:-/usr/local/nvidia/lib64}"
ENTRY="${TRANSPORT_ENTRY:-/scripts/container_entry.sh}"
if [ -x "${ENTRY}" ]; then "${ENTRY}" install --install-nccl; fi
mkdir -p "${DST}"
cp -r "${SRC}/." "${DST}"
if [ -n "${TRANSPORT_EXTRA_SRC:-}" ] && [ -n "${TRANSPORT_EXTRA_DST:-}" ]; then   # gIB also ships its whole tree (scripts, tuner configs)
  mkdir -p "${TRANSPORT_EXTRA_DST}"
  cp -r "${TRANSPORT_EXTRA_SRC}/." "${TRANSPORT_EXTRA_DST}"
fi
echo "installed transport payload from ${SRC} into ${DST}"
