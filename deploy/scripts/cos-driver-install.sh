#!/bin/bash
# COS driver install step (initContainer of the driver-installer DaemonSets).
# Skips when an nvidia module is already loaded, else runs cos-gpu-installer and opens up the install dir.
# Behaviour: reference nvidia-driver-installer/cos/daemonset-preloaded-latest.yaml:112-124.
set -u
VERSION_FLAG="${COS_GPU_INSTALLER_VERSION_FLAG:---version=latest}"
EXTRA_FLAGS="${COS_GPU_INSTALLER_EXTRA_FLAGS:-}"
ROOT="${ROOT_MOUNT_DIR:-/root}"
INSTALLER="${COS_GPU_INSTALLER:-/cos-gpu-installer}"
echo "Checking for existing GPU driver modules"
if ${LSMOD:-lsmod} | grep -q nvidia; then
  echo "GPU driver is already installed; the loaded version may differ from the one requested, skipping installation"
  exit 0
fi
echo "No GPU driver module detected, installing now"
# shellcheck disable=SC2086
"${INSTALLER}" install ${VERSION_FLAG} ${EXTRA_FLAGS} || exit 1
chmod 755 "${ROOT}/home/kubernetes/bin/nvidia"
