#!/bin/bash
# minikube host: no header packages exist for its kernel, so fetch the matching source from cdn.kernel.org, reuse the
# running kernel's config, `make modules_prepare`, and point the NVIDIA installer at it
# (reference nvidia-driver-installer/minikube/entrypoint.sh:34-111,205-220).
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
# shellcheck disable=SC1091
. "${DRIVER_INSTALL_LIB:-${HERE}/../lib/driver-install-lib.sh}"

KERNEL_SRC_DIR="${KERNEL_SRC_DIR:-/usr/src/linux}"
export KERNEL_SOURCE_PATH="${KERNEL_SRC_DIR}"

# "4.19.0" is published as linux-4.19.tar.xz: drop a trailing ".0" patch level, and any local suffix ("-minikube").
kernel_tarball_version() {
  local v="${KERNEL_VERSION%%-*}"
  if [[ "${v}" =~ ^([0-9]+\.[0-9]+)\.0$ ]]; then v="${BASH_REMATCH[1]}"; fi
  echo "${v}"
}

download_kernel_src() {
  local v major; v="$(kernel_tarball_version)"; major="${v%%.*}"
  say "kernel sources: linux-${v} from cdn.kernel.org"
  mkdir -p "${KERNEL_SRC_DIR}"
  ${CURL} -L -S -f "https://cdn.kernel.org/pub/linux/kernel/v${major}.x/linux-${v}.tar.xz" -o /tmp/linux.tar.xz || return 1
  ${TAR:-tar} -xf /tmp/linux.tar.xz -C "${KERNEL_SRC_DIR}" --strip-components=1 || return 1
  if [[ -r /proc/config.gz ]]; then zcat /proc/config.gz > "${KERNEL_SRC_DIR}/.config"; fi
  ( cd "${KERNEL_SRC_DIR}" && ${MAKE:-make} olddefconfig && ${MAKE:-make} modules_prepare ) || return 1
  # the module's vermagic must match the running kernel including its local suffix
  echo "#define UTS_RELEASE \"${KERNEL_VERSION}\"" > "${KERNEL_SRC_DIR}/include/generated/utsrelease.h"
}

if [[ "${BASH_SOURCE[0]}" == "$0" ]]; then install_driver_main download_kernel_src; fi
