#!/bin/bash
# `kubectl exec host-1 -- /scripts/allgather.sh host1 host2`: set up ssh + hostfiles and run all_gather_perf over both pods.
# Behaviour: reference gpudirect-tcpx/nccl-config.yaml:6-17.
set -eu
/scripts/init_ssh.sh "$@"
pushd /scripts >/dev/null
/scripts/gen_hostfiles.sh "$@"
popd >/dev/null
/scripts/run-nccl.sh "${BENCHMARK:-all_gather_perf}" "${LD_LIBRARY_PATH}" 8 eth1,eth2,eth3,eth4 "${MIN_BYTES:-1M}" "${MAX_BYTES:-512M}" "$#"
