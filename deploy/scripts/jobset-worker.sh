#!/bin/bash
# JobSet worker for multi-node nccl-tests: start sshd, rank 0 waits until every peer answers ssh, writes the hostfile
# (slots = GPUs per node) and runs the benchmark through mpirun; the others wait for the head to finish.
# Behaviour: reference gpudirect-rdma/nccl-test-a4x-max-jobset.yaml:104-163.
set -u
NUM_NODES="${NUM_NODES:?}"; SLOTS="${GPUS_PER_NODE:-4}"; BENCH="${BENCHMARK:-all_gather_perf}"
JOB="${JOBSET_NAME:?}"; RJ="${REPLICATED_JOB_NAME:-w}"; SSH_PORT="${SSH_PORT:-222}"
service ssh restart 2>/dev/null || /usr/sbin/sshd -p "${SSH_PORT}"
idx="${JOB_COMPLETION_INDEX:-0}"
host() { echo "${JOB}-${RJ}-0-$1.${JOB}"; }
if [ "${idx}" != "0" ]; then
  until ssh -p "${SSH_PORT}" -o StrictHostKeyChecking=no "$(host 0)" true 2>/dev/null; do sleep 5; done     # head is up
  while ssh -p "${SSH_PORT}" -o StrictHostKeyChecking=no "$(host 0)" pgrep -f mpirun >/dev/null 2>&1 || [ ! -f /tmp/head-started ]; do touch /tmp/head-started; sleep 10; done
  exit 0
fi
: > /tmp/hostfile
for i in $(seq 0 $(( NUM_NODES - 1 ))); do
  until ssh -p "${SSH_PORT}" -o StrictHostKeyChecking=no "$(host "$i")" true 2>/dev/null; do echo "waiting for $(host "$i")"; sleep 5; done
  echo "$(host "$i") slots=${SLOTS}" >> /tmp/hostfile
done
# shellcheck disable=SC1091
source /usr/local/gib/scripts/set_nccl_env.sh
mpirun --allow-run-as-root --hostfile /tmp/hostfile -np $(( NUM_NODES * SLOTS )) --mca plm_rsh_args "-p ${SSH_PORT}" --mca btl tcp,self --mca btl_tcp_if_include eth0 \
  -x LD_LIBRARY_PATH -x NCCL_TESTS_SPLIT_MASK="${NCCL_TESTS_SPLIT_MASK:-0x0}" $(env | grep -E '^NCCL_' | cut -d= -f1 | sed 's/^/-x /') \
  "/third_party/nccl-tests/build/${BENCH}" -b 1K -e 8G -f 2 -g 1 -w 5 --iters 100 -c 1
