#!/bin/bash
# compute-sanitizer over the collective kernels with virtual ranks on one GPU (cross-GPU flag protocols are
# where races hide). memcheck + synccheck on small sizes; every run under its own timeout (the tools slow spinning kernels a lot).
mkdir -p gpurun_out; export B200COLL_TIMEOUT_MS=20000
P=./build/b200coll_perf
: > gpurun_out/sanitizer.log
for tool in memcheck synccheck; do
  for spec in "all_reduce ll" "all_reduce ll2" "all_reduce twoshot" "all_gather ll" "reduce_scatter twoshot" "alltoall twoshot"; do
    set -- $spec
    echo "=== $tool $1 $2" >> gpurun_out/sanitizer.log
    timeout 120 compute-sanitizer --tool $tool --error-exitcode 9 $P --devs 0,0 --op $1 --algo $2 -b 4K -e 64K -f 4 --iters 2 --warmup 1 >> gpurun_out/sanitizer.log 2>&1
    echo "rc=$?" >> gpurun_out/sanitizer.log
  done
done
grep -E "^===|^rc=|ERROR SUMMARY" gpurun_out/sanitizer.log
