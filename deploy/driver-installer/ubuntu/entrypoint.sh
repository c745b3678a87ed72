#!/bin/bash
# Ubuntu host: kernel headers come from apt (reference nvidia-driver-installer/ubuntu/entrypoint.sh:70-74,165-178).
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
# shellcheck disable=SC1091
. "${DRIVER_INSTALL_LIB:-${HERE}/../lib/driver-install-lib.sh}"

download_kernel_headers() {
  ${APT_GET:-apt-get} update && ${APT_GET:-apt-get} install -y "linux-headers-${KERNEL_VERSION}"
}

if [[ "${BASH_SOURCE[0]}" == "$0" ]]; then install_driver_main download_kernel_headers; fi
