#!/bin/bash
# Source me in the workload container (the nccl-env-profile.sh / set_nccl_env.sh analogue; reference
# gpudirect-tcpxo/README.md:25-35, gpudirect-rdma/nccl-test-a4.yaml:76). The device plugin exports the same values from
# Allocate when GPUConfig.Transport is "b200coll", so this file is only needed for pods that bypass the plugin.
export B200COLL_LIB_DIR="${B200COLL_LIB_DIR:-/usr/local/nvidia/lib64}"
export B200COLL_LIB="${B200COLL_LIB_DIR}/libb200coll.so"
export LD_LIBRARY_PATH="${B200COLL_LIB_DIR}${LD_LIBRARY_PATH:+:${LD_LIBRARY_PATH}}"
export B200COLL_NVLS="${B200COLL_NVLS:--1}"          # -1 probe, 0 P2P only, 1 require multicast
export B200COLL_ALGO="${B200COLL_ALGO:-auto}"         # auto | ll | ll2 | oneshot | twoshot | nvls
export B200COLL_TIMEOUT_MS="${B200COLL_TIMEOUT_MS:-600000}"
export B200COLL_DEBUG="${B200COLL_DEBUG:-WARN}"
if [ -f "${B200COLL_LIB_DIR}/b200_nvswitch.tbl" ]; then export B200COLL_TUNER_FILE="${B200COLL_TUNER_FILE:-${B200COLL_LIB_DIR}/b200_nvswitch.tbl}"; fi
