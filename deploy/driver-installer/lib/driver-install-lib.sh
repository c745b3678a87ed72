#!/bin/bash
# Shared steps of the NVIDIA .run driver installers (Ubuntu and minikube hosts).
#
# Flow (reference nvidia-driver-installer/ubuntu/entrypoint.sh:33-178, minikube/entrypoint.sh:34-220):
#   cache hit (same kernel + driver version recorded in <install dir>/.cache) -> insmod the cached modules, verify
#   else: get kernel headers/sources -> redirect the installer's hard-wired output dirs (/usr/bin,
#         /usr/lib/x86_64-linux-gnu, /lib/modules/$K/video) into the host-visible install dir with overlayfs ->
#         download + run the .run file -> record the cache -> verify (nvidia-smi, nvidia-modprobe -c0 -u)
#   finally add <host install dir>/lib64 to the host's ld.so.conf and refresh its cache.
# Differences from the reference: one table drives the three overlay redirects; Blackwell needs the open kernel
# modules, so NVIDIA_KERNEL_MODULE_TYPE defaults to "open" for driver >= 560; every external command can be replaced
# through an env var so the script is unit-testable without root.
set -o pipefail
set -u

NVIDIA_DRIVER_BRANCH="${NVIDIA_DRIVER_BRANCH:-tesla}"
NVIDIA_DRIVER_VERSION="${NVIDIA_DRIVER_VERSION:-570.124.06}"
NVIDIA_DRIVER_DOWNLOAD_URL="${NVIDIA_DRIVER_DOWNLOAD_URL:-https://us.download.nvidia.com/${NVIDIA_DRIVER_BRANCH}/${NVIDIA_DRIVER_VERSION}/NVIDIA-Linux-x86_64-${NVIDIA_DRIVER_VERSION}.run}"
NVIDIA_INSTALL_DIR_HOST="${NVIDIA_INSTALL_DIR_HOST:-/home/kubernetes/bin/nvidia}"
NVIDIA_INSTALL_DIR_CONTAINER="${NVIDIA_INSTALL_DIR_CONTAINER:-/usr/local/nvidia}"
NVIDIA_INSTALLER_RUNFILE="$(basename "${NVIDIA_DRIVER_DOWNLOAD_URL}")"
ROOT_MOUNT_DIR="${ROOT_MOUNT_DIR:-/root}"
CACHE_FILE="${NVIDIA_INSTALL_DIR_CONTAINER}/.cache"
KERNEL_VERSION="${KERNEL_VERSION:-$(uname -r)}"
LD_SO_CONF_D="${LD_SO_CONF_D:-/etc/ld.so.conf.d}"
# command seams
: "${MOUNT:=mount}" "${UMOUNT:=umount}" "${LDCONFIG:=ldconfig}" "${LSMOD:=lsmod}" "${INSMOD:=insmod}" "${CURL:=curl}"

driver_major() { echo "${NVIDIA_DRIVER_VERSION%%.*}"; }

kernel_module_type_flag() {
  local t="${NVIDIA_KERNEL_MODULE_TYPE:-}"
  if [[ -z "${t}" && "$(driver_major)" -ge 560 ]]; then t=open; fi
  if [[ -n "${t}" ]]; then echo "--kernel-module-type=${t}"; fi
}

check_cached_version() {
  echo "Checking cached version"
  if [[ ! -f "${CACHE_FILE}" ]]; then echo "Cache file ${CACHE_FILE} not found."; return 1; fi
  local CACHE_KERNEL_VERSION="" CACHE_NVIDIA_DRIVER_VERSION=""
  # shellcheck disable=SC1090
  . "${CACHE_FILE}"
  if [[ "${KERNEL_VERSION}" == "${CACHE_KERNEL_VERSION}" && "${NVIDIA_DRIVER_VERSION}" == "${CACHE_NVIDIA_DRIVER_VERSION}" ]]; then
    echo "Found existing driver installation for kernel version ${KERNEL_VERSION} and driver version ${NVIDIA_DRIVER_VERSION}."
    return 0
  fi
  echo "Cache file ${CACHE_FILE} found but existing versions didn't match."
  return 1
}

update_cached_version() {
  printf 'CACHE_KERNEL_VERSION=%s\nCACHE_NVIDIA_DRIVER_VERSION=%s\n' "${KERNEL_VERSION}" "${NVIDIA_DRIVER_VERSION}" > "${CACHE_FILE}"
  echo "Updated cached version as:"; cat "${CACHE_FILE}"
}

update_container_ld_cache() {
  echo "${NVIDIA_INSTALL_DIR_CONTAINER}/lib64" > "${LD_SO_CONF_D}/nvidia.conf"
  ${LDCONFIG}
}

# installer-owned dir | subdir of the install dir that must receive its contents
overlay_table() {
  cat <<__T__
/usr/bin|bin
/usr/lib/x86_64-linux-gnu|lib64
/lib/modules/${KERNEL_VERSION}/video|drivers
__T__
}

configure_nvidia_installation_dirs() {
  echo "Configuring installation directories..."
  mkdir -p "${NVIDIA_INSTALL_DIR_CONTAINER}"
  local mounted=()
  while IFS='|' read -r lower sub; do
    mkdir -p "${NVIDIA_INSTALL_DIR_CONTAINER}/${sub}" "${NVIDIA_INSTALL_DIR_CONTAINER}/${sub}-workdir" "${OVERLAY_ROOT:-}${lower}"
    ${MOUNT} -t overlay -o "lowerdir=${lower},upperdir=${NVIDIA_INSTALL_DIR_CONTAINER}/${sub},workdir=${NVIDIA_INSTALL_DIR_CONTAINER}/${sub}-workdir" none "${lower}" || return 1
    mounted=("${lower}" "${mounted[@]}")
  done < <(overlay_table)
  update_container_ld_cache          # keeps nvidia-installer's log free of ldconfig warnings
  # shellcheck disable=SC2064
  trap "for m in ${mounted[*]}; do ${UMOUNT} \$m; done" EXIT
  echo "Configuring installation directories... DONE."
}

download_nvidia_installer() {
  echo "Downloading Nvidia installer..."
  ${CURL} -L -S -f "${NVIDIA_DRIVER_DOWNLOAD_URL}" -o "${NVIDIA_INSTALL_DIR_CONTAINER}/${NVIDIA_INSTALLER_RUNFILE}" || return 1
}

run_nvidia_installer() {
  echo "Running Nvidia installer..."
  local extra=()
  [[ -n "${KERNEL_SOURCE_PATH:-}" ]] && extra+=("--kernel-source-path=${KERNEL_SOURCE_PATH}")
  local kmt; kmt="$(kernel_module_type_flag)"; [[ -n "${kmt}" ]] && extra+=("${kmt}")
  ( cd "${NVIDIA_INSTALL_DIR_CONTAINER}" && ${SH:-sh} "${NVIDIA_INSTALLER_RUNFILE}" \
      --utility-prefix="${NVIDIA_INSTALL_DIR_CONTAINER}" --opengl-prefix="${NVIDIA_INSTALL_DIR_CONTAINER}" --no-install-compat32-libs \
      --log-file-name="${NVIDIA_INSTALL_DIR_CONTAINER}/nvidia-installer.log" --no-drm --silent --accept-license "${extra[@]}" ) || return 1
  echo "Running Nvidia installer... DONE."
}

configure_cached_installation() {
  echo "Configuring cached driver installation..."
  update_container_ld_cache
  if ! ${LSMOD} | grep -qw nvidia; then ${INSMOD} "${NVIDIA_INSTALL_DIR_CONTAINER}/drivers/nvidia.ko" || return 1; fi
  if ! ${LSMOD} | grep -qw nvidia_uvm; then ${INSMOD} "${NVIDIA_INSTALL_DIR_CONTAINER}/drivers/nvidia-uvm.ko" || return 1; fi
}

verify_nvidia_installation() {
  echo "Verifying Nvidia installation..."
  export PATH="${NVIDIA_INSTALL_DIR_CONTAINER}/bin:${PATH}"
  nvidia-smi || return 1
  nvidia-modprobe -c0 -u || return 1      # creates /dev/nvidia-uvm
}

update_host_ld_cache() {
  echo "Updating host's ld cache..."
  local conf="${ROOT_MOUNT_DIR}/etc/ld.so.conf"
  grep -qxF "${NVIDIA_INSTALL_DIR_HOST}/lib64" "${conf}" 2>/dev/null || echo "${NVIDIA_INSTALL_DIR_HOST}/lib64" >> "${conf}"     # idempotent (the reference appends on every run)
  ${LDCONFIG} -r "${ROOT_MOUNT_DIR}"
}

install_driver_main() {   # $1 = function that fetches kernel headers/sources for this distro
  local fetch_kernel="$1"
  if check_cached_version; then
    configure_cached_installation && verify_nvidia_installation || return 1
  else
    "${fetch_kernel}" && configure_nvidia_installation_dirs && download_nvidia_installer && run_nvidia_installer && update_cached_version && verify_nvidia_installation || return 1
  fi
  update_host_ld_cache
}
