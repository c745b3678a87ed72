#!/bin/bash
# Confidential-GPU variant: record the confidential node type for the persistenced sidecar, install without signature
# verification, insmod the four modules by hand and create the device nodes.
# Behaviour: reference nvidia-driver-installer/cos/daemonset-confidential-latest.yaml:117-143.
set -u
ROOT="${ROOT_MOUNT_DIR:-/root}"
MD="${METADATA_URL:-http://metadata.google.internal/computeMetadata/v1}"
NVIDIA_DIR="${NVIDIA_INSTALL_DIR_CONTAINER:-/usr/local/nvidia}"
INSTALLER="${COS_GPU_INSTALLER:-/cos-gpu-installer}"
labels=$(${CURL:-curl} -sf -H "Metadata-Flavor: Google" "${MD}/instance/attributes/kube-labels" || true)
node_type=$(echo "${labels}" | tr ',' '\n' | sed -n 's/^cloud.google.com\/gke-confidential-nodes-instance-type=//p' | head -n1)
mkdir -p "${ROOT}/etc/nvidia"
echo "${node_type}" > "${ROOT}/etc/nvidia/confidential_node_type.txt"
echo "confidential node type: ${node_type:-<none>}"
if ${LSMOD:-lsmod} | grep -q nvidia; then
  echo "GPU driver is already installed, skipping installation"
  exit 0
fi
"${INSTALLER}" install ${COS_GPU_INSTALLER_VERSION_FLAG:---version=latest} --no-verify || exit 1
chmod 755 "${ROOT}/home/kubernetes/bin/nvidia"
for mod in nvidia nvidia-uvm nvidia-modeset nvidia-drm; do
  ${INSMOD:-insmod} "${NVIDIA_DIR}/drivers/${mod}.ko" || echo "insmod ${mod} failed (may already be loaded)"
done
"${NVIDIA_DIR}/bin/nvidia-modprobe" -c0 -u -m || exit 1
