#!/bin/bash
# NVIDIA .run driver installation for hosts without a vendor driver image (Ubuntu and minikube nodes): the part the two
# entrypoints share. Role parity: nvidia-driver-installer/ubuntu/entrypoint.sh and minikube/entrypoint.sh of the reference
# (SURVEY C17/C18). What is fixed by that role — and therefore identical here — is the environment contract (the NVIDIA_* / ROOT_MOUNT_DIR
# variables the DaemonSets set), the two-line stamp file other tooling reads, and the nvidia-installer command line. Everything else
# is organised differently:
#
#   * the work is a PLAN: an ordered list of step functions chosen up front ("reuse" when the stamp matches this kernel + driver,
#     "build" otherwise) and executed by one runner that numbers the steps, times them and stops at the first failure;
#   * the stamp is PARSED (two known keys), never sourced as shell;
#   * the installer's three hard-wired output directories are redirected into the host-visible install dir from one table;
#   * Blackwell needs the open kernel modules: --kernel-module-type=open is passed for driver branches >= 560 unless
#     NVIDIA_KERNEL_MODULE_TYPE says otherwise;
#   * every external program is reached through a variable (MOUNT, CURL, ...) so the whole flow runs against stubs in tests/test_manifests.py;
#   * publishing the library directory to the host's loader configuration is idempotent.
set -o pipefail
set -u

NVIDIA_DRIVER_BRANCH="${NVIDIA_DRIVER_BRANCH:-tesla}"
NVIDIA_DRIVER_VERSION="${NVIDIA_DRIVER_VERSION:-570.124.06}"
NVIDIA_DRIVER_DOWNLOAD_URL="${NVIDIA_DRIVER_DOWNLOAD_URL:-https://us.download.nvidia.com/${NVIDIA_DRIVER_BRANCH}/${NVIDIA_DRIVER_VERSION}/NVIDIA-Linux-x86_64-${NVIDIA_DRIVER_VERSION}.run}"
NVIDIA_INSTALL_DIR_HOST="${NVIDIA_INSTALL_DIR_HOST:-/home/kubernetes/bin/nvidia}"
NVIDIA_INSTALL_DIR_CONTAINER="${NVIDIA_INSTALL_DIR_CONTAINER:-/usr/local/nvidia}"
ROOT_MOUNT_DIR="${ROOT_MOUNT_DIR:-/root}"
KERNEL_VERSION="${KERNEL_VERSION:-$(uname -r)}"
LD_SO_CONF_D="${LD_SO_CONF_D:-/etc/ld.so.conf.d}"
: "${MOUNT:=mount}" "${UMOUNT:=umount}" "${LDCONFIG:=ldconfig}" "${LSMOD:=lsmod}" "${INSMOD:=insmod}" "${CURL:=curl}"

DRV_PREFIX="${NVIDIA_INSTALL_DIR_CONTAINER}"
DRV_STAMP="${DRV_PREFIX}/.cache"
DRV_RUNFILE="${NVIDIA_DRIVER_DOWNLOAD_URL##*/}"
DRV_OVERLAYS=()            # mount points created by step_redirect_outputs, newest first

say() { printf '[driver-install] %s\n' "$*"; }

# ---- stamp: which (kernel, driver) pair the modules under ${DRV_PREFIX}/drivers were built for
stamp_value() {            # $1 = key; prints the value recorded in the stamp, nothing if absent
  [[ -r "${DRV_STAMP}" ]] || return 0
  local line
  while IFS= read -r line; do
    if [[ "${line}" == "$1="* ]]; then printf '%s' "${line#*=}"; return 0; fi
  done < "${DRV_STAMP}"
}

stamp_matches() {
  [[ -f "${DRV_STAMP}" ]] || { say "no stamp at ${DRV_STAMP}: nothing has been installed here yet"; return 1; }
  local k d; k="$(stamp_value CACHE_KERNEL_VERSION)"; d="$(stamp_value CACHE_NVIDIA_DRIVER_VERSION)"
  if [[ "${k}" == "${KERNEL_VERSION}" && "${d}" == "${NVIDIA_DRIVER_VERSION}" ]]; then
    say "stamp matches: driver ${d} was built for kernel ${k}; reusing it"
    return 0
  fi
  say "stamp is for kernel '${k}' / driver '${d}', wanted '${KERNEL_VERSION}' / '${NVIDIA_DRIVER_VERSION}': rebuilding"
  return 1
}

step_write_stamp() {
  printf 'CACHE_KERNEL_VERSION=%s\nCACHE_NVIDIA_DRIVER_VERSION=%s\n' "${KERNEL_VERSION}" "${NVIDIA_DRIVER_VERSION}" > "${DRV_STAMP}"
}

# ---- loader configuration
refresh_container_loader() {
  printf '%s\n' "${DRV_PREFIX}/lib64" > "${LD_SO_CONF_D}/nvidia.conf"
  ${LDCONFIG}
}

step_publish_to_host_loader() {
  local conf="${ROOT_MOUNT_DIR}/etc/ld.so.conf" want="${NVIDIA_INSTALL_DIR_HOST}/lib64"
  if ! grep -qxF "${want}" "${conf}" 2>/dev/null; then printf '%s\n' "${want}" >> "${conf}"; fi
  ${LDCONFIG} -r "${ROOT_MOUNT_DIR}"
}

# ---- build path
open_modules_flag() {      # prints the nvidia-installer flag selecting the kernel-module flavour, if one is needed
  local flavour="${NVIDIA_KERNEL_MODULE_TYPE:-}"
  if [[ -z "${flavour}" ]] && (( ${NVIDIA_DRIVER_VERSION%%.*} >= 560 )); then flavour=open; fi
  [[ -n "${flavour}" ]] && printf -- '--kernel-module-type=%s' "${flavour}"
  return 0
}

# where nvidia-installer insists on writing  ->  directory under the install prefix that must end up holding it
redirect_table() {
  printf '%s %s\n' /usr/bin bin /usr/lib/x86_64-linux-gnu lib64 "/lib/modules/${KERNEL_VERSION}/video" drivers
}

undo_redirects() {
  local m
  for m in "${DRV_OVERLAYS[@]}"; do ${UMOUNT} "${m}"; done
  DRV_OVERLAYS=()
}

step_redirect_outputs() {
  mkdir -p "${DRV_PREFIX}"
  trap undo_redirects EXIT
  local target sub upper
  while read -r target sub; do
    upper="${DRV_PREFIX}/${sub}"
    mkdir -p "${upper}" "${upper}-workdir" "${OVERLAY_ROOT:-}${target}"
    ${MOUNT} -t overlay -o "lowerdir=${target},upperdir=${upper},workdir=${upper}-workdir" none "${target}" || return 1
    DRV_OVERLAYS=("${target}" "${DRV_OVERLAYS[@]}")
  done < <(redirect_table)
  refresh_container_loader          # so that nvidia-installer's own ldconfig run finds the redirected lib64
}

step_fetch_runfile() {
  ${CURL} -L -S -f "${NVIDIA_DRIVER_DOWNLOAD_URL}" -o "${DRV_PREFIX}/${DRV_RUNFILE}"
}

step_run_runfile() {
  local args=(--utility-prefix="${DRV_PREFIX}" --opengl-prefix="${DRV_PREFIX}" --no-install-compat32-libs
              --log-file-name="${DRV_PREFIX}/nvidia-installer.log" --no-drm --silent --accept-license)
  [[ -n "${KERNEL_SOURCE_PATH:-}" ]] && args+=("--kernel-source-path=${KERNEL_SOURCE_PATH}")
  local flavour; flavour="$(open_modules_flag)"
  [[ -n "${flavour}" ]] && args+=("${flavour}")
  ( cd "${DRV_PREFIX}" && ${SH:-sh} "${DRV_RUNFILE}" "${args[@]}" )
}

# ---- reuse path
step_load_prebuilt_modules() {
  refresh_container_loader
  local mod
  for mod in nvidia nvidia-uvm; do
    if ! ${LSMOD} | grep -qw "${mod//-/_}"; then ${INSMOD} "${DRV_PREFIX}/drivers/${mod}.ko" || return 1; fi
  done
}

# ---- both paths
step_check_driver_answers() {
  PATH="${DRV_PREFIX}/bin:${PATH}" nvidia-smi || return 1
  PATH="${DRV_PREFIX}/bin:${PATH}" nvidia-modprobe -c0 -u          # also creates /dev/nvidia-uvm
}

run_plan() {               # run the named step functions in order; stop at the first one that fails
  local total=$# n=0 step t0
  for step in "$@"; do
    n=$((n + 1)); t0=${SECONDS}
    say "(${n}/${total}) ${step#step_}"
    if ! "${step}"; then say "(${n}/${total}) ${step#step_} FAILED after $((SECONDS - t0)) s"; return 1; fi
  done
}

install_driver_main() {    # $1 = the distro's step that provides kernel headers or sources
  local kernel_step="$1"
  say "driver ${NVIDIA_DRIVER_VERSION} (${NVIDIA_DRIVER_BRANCH}) for kernel ${KERNEL_VERSION}; prefix ${DRV_PREFIX} (host: ${NVIDIA_INSTALL_DIR_HOST})"
  if stamp_matches; then
    run_plan step_load_prebuilt_modules step_check_driver_answers step_publish_to_host_loader
  else
    run_plan "${kernel_step}" step_redirect_outputs step_fetch_runfile step_run_runfile step_write_stamp step_check_driver_answers step_publish_to_host_loader
  fi
}
