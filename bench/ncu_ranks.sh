#!/bin/bash
# One ncu instance PER RANK (never one ncu around a multi-rank launcher): every rank process of an nccl-tests-style run is started under
# its own `ncu` with a metric list small enough for a single pass, so no kernel is ever replayed and the cross-rank flag protocols see
# each launch exactly once. ncu still serialises the kernels of its own process, which is how the ranks of one-rank-per-GPU jobs run anyway.
#   bench/ncu_ranks.sh <nranks> <tag> <kernel-regex> <metrics> <perf binary and flags ...>
# Output: gpurun_out/ncu_<tag>_r<rank>.csv (ncu --csv), gpurun_out/ncu_<tag>_r<rank>.out (the tool's own table).
N=$1; TAG=$2; KRE=$3; METRICS=$4; shift 4
mkdir -p gpurun_out
PORT=$((31000 + RANDOM % 2000))
for r in $(seq 0 $((N - 1))); do
  RANK=$r WORLD_SIZE=$N LOCAL_RANK=$r MASTER_ADDR=127.0.0.1 MASTER_PORT=$PORT B200COLL_JOB_ID=ncu-$TAG-$PORT B200COLL_TIMEOUT_MS=${B200COLL_TIMEOUT_MS:-8000} \
    timeout ${NCU_TIMEOUT:-150} ncu --metrics "$METRICS" --clock-control none --cache-control none -k regex:"$KRE" --launch-skip ${NCU_SKIP:-4} -c ${NCU_COUNT:-2} \
    --csv --log-file gpurun_out/ncu_${TAG}_r$r.csv "$@" > gpurun_out/ncu_${TAG}_r$r.out 2>&1 &
done
wait
head -c 600 gpurun_out/ncu_${TAG}_r0.csv | tail -c 300; echo
