#!/bin/bash
# Write the GCE machine type where the persistenced/gridd sidecar looks for it.
# Behaviour: reference nvidia-driver-installer/cos/daemonset-vgpu-latest.yaml:117-157.
set -u
ROOT="${ROOT_MOUNT_DIR:-/root}"
MD="${METADATA_URL:-http://metadata.google.internal/computeMetadata/v1}"
mt=$(${CURL:-curl} -sf -H "Metadata-Flavor: Google" "${MD}/instance/machine-type" || true)
mkdir -p "${ROOT}/etc/nvidia"
basename "${mt}" > "${ROOT}/etc/nvidia/machine_type.txt"
echo "machine type: $(cat "${ROOT}/etc/nvidia/machine_type.txt")"
