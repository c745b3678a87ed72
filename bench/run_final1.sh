#!/bin/bash
# One-GPU acceptance pass for the tree as committed: the GPU test-suite, the driver's smoke entry point, and a
# short default bench run. Everything is wrapped in `timeout` so a hang costs seconds, not the box.
mkdir -p gpurun_out
echo "== $(date -u +%T) pytest -m gpu"
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/pytest_gpu_final.log
echo "== $(date -u +%T) smoke"
timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"
tail -n 2 gpurun_out/smoke_final.log
echo "== $(date -u +%T) rooted ops, 4 virtual ranks"
for op in broadcast reduce; do timeout 60 build/b200coll_perf --devs 0,0,0,0 --op $op -b 1K -e 16M -f 8 --iters 5 --warmup 2 > gpurun_out/rooted_$op.txt 2>&1; echo "$op rc=$?"; tail -n 4 gpurun_out/rooted_$op.txt; done
echo "== $(date -u +%T) done"
