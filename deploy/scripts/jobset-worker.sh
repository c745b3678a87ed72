#!/bin/bash
# JobSet worker for multi-node nccl-tests: start sshd; rank 0 waits until every peer answers ssh, writes the hostfile
# (slots = GPUs per node) and runs the benchmark through mpirun; every other rank stays up for as long as the head's
# container answers ssh (that is: until mpirun has finished there) and then exits 0 so the JobSet completes.
# Behaviour: reference gpudirect-rdma/nccl-test-a4x-max-jobset.yaml:104-163.
set -u
NUM_NODES="${NUM_NODES:?}"; SLOTS="${GPUS_PER_NODE:-4}"; BENCH="${BENCHMARK:-all_gather_perf}"
JOB="${JOBSET_NAME:?}"; RJ="${REPLICATED_JOB_NAME:-w}"; SSH_PORT="${SSH_PORT:-222}"
SSH="${SSH:-ssh}"; MPIRUN="${MPIRUN:-mpirun}"; HOSTFILE="${HOSTFILE:-/tmp/hostfile}"; POLL_S="${POLL_S:-5}"
NCCL_ENV_SCRIPT="${NCCL_ENV_SCRIPT:-/usr/local/gib/scripts/set_nccl_env.sh}"
if [ -n "${SSHD_START:-}" ]; then ${SSHD_START}; else service ssh restart 2>/dev/null || /usr/sbin/sshd -p "${SSH_PORT}"; fi
idx="${JOB_COMPLETION_INDEX:-0}"
host() { echo "${JOB}-${RJ}-0-$1.${JOB}"; }
reach() { ${SSH} -p "${SSH_PORT}" -o StrictHostKeyChecking=no "$1" true 2>/dev/null; }
if [ "${idx}" != "0" ]; then
  until reach "$(host 0)"; do sleep "${POLL_S}"; done          # the head is up
  while reach "$(host 0)"; do sleep "${POLL_S}"; done          # ... and now it is gone: the benchmark is over
  exit 0
fi
: > "${HOSTFILE}"
for i in $(seq 0 $(( NUM_NODES - 1 ))); do
  until reach "$(host "$i")"; do echo "waiting for $(host "$i")"; sleep "${POLL_S}"; done
  echo "$(host "$i") slots=${SLOTS}" >> "${HOSTFILE}"
done
# shellcheck disable=SC1090
if [ -f "${NCCL_ENV_SCRIPT}" ]; then source "${NCCL_ENV_SCRIPT}"; fi
# every NCCL_* variable of this shell travels to the ranks
${MPIRUN} --allow-run-as-root --hostfile "${HOSTFILE}" -np $(( NUM_NODES * SLOTS )) --mca plm_rsh_args "-p ${SSH_PORT}" --mca btl tcp,self --mca btl_tcp_if_include eth0 \
  -x LD_LIBRARY_PATH -x NCCL_TESTS_SPLIT_MASK="${NCCL_TESTS_SPLIT_MASK:-0x0}" $(env | grep -E '^NCCL_' | cut -d= -f1 | grep -v '^NCCL_TESTS_SPLIT_MASK$' | sed 's/^/-x /') \
  "${NCCL_TESTS_DIR:-/third_party/nccl-tests/build}/${BENCH}" -b 1K -e 8G -f 2 -g 1 -w 5 --iters 100 -c 1
