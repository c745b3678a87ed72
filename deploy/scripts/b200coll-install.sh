#!/bin/bash
# b200coll transport installer: drop libb200coll.so (+ the NCCL-API shim, the perf tool and the env profile) into the
# host lib dir, then run the start-up self check so a box that cannot run the fast paths is reported at install time
# (the guest-config-checker role, reference gpudirect-tcpxo/README.md:84,239).
set -eu
SRC="${B200COLL_SRC_DIR:-/opt/b200coll}"
DST="${NCCL_INSTALL_DIR:-/usr/local/nvidia/lib64}"
BIN="${B200COLL_BIN_DIR:-/usr/local/nvidia/bin}"
mkdir -p "${DST}" "${BIN}"
cp "${SRC}/lib/libb200coll.so" "${SRC}/lib/libb200coll_nccl.so" "${DST}/"
cp "${SRC}/bin/b200coll_perf" "${BIN}/"
if [ -f "${SRC}/bin/mps_probe" ]; then cp "${SRC}/bin/mps_probe" "${BIN}/"; fi      # deploy/example/cuda-mps/mps-probe.yaml runs it from the host mount
# nccl-tests names, so `run-nccl.sh all_gather_perf ...` style scripts keep working (the binary dispatches on argv[0])
for n in all_reduce_perf all_gather_perf reduce_scatter_perf alltoall_perf broadcast_perf reduce_perf sendrecv_perf gather_perf scatter_perf hypercube_perf; do ln -sf b200coll_perf "${BIN}/${n}"; done
cp "${SRC}/b200coll-env-profile.sh" "${DST}/"
if [ -f "${SRC}/tuner/b200_nvswitch.tbl" ]; then cp "${SRC}/tuner/b200_nvswitch.tbl" "${DST}/"; fi
if [ "${B200COLL_SKIP_SELFCHECK:-0}" != "1" ]; then
  LD_LIBRARY_PATH="${DST}:${LD_LIBRARY_PATH:-}" "${BIN}/b200coll_perf" --selfcheck || { echo "b200coll self-check failed: this node cannot run the NVLink fast paths"; exit 1; }
fi
echo "installed libb200coll into ${DST}"
