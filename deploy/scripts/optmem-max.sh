#!/bin/bash
# devmem-TCP needs a larger ancillary-buffer limit (reference gpudirect-tcpx/optmem-max-ds.yaml:35-38).
set -eu
echo "${OPTMEM_MAX:-131072}" > "${PROC_SYS:-/proc/sys}/net/core/optmem_max"
echo "optmem_max set to $(cat "${PROC_SYS:-/proc/sys}/net/core/optmem_max")"
