#!/bin/bash
# 1-GPU sanity of the shipped tree + the opt-in PDL path (virtual ranks share cuda:0).
mkdir -p gpurun_out; export B200COLL_TIMEOUT_MS=5000
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log
B200COLL_PDL=1 timeout 300 python -m pytest tests/test_coll_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu_pdl.log 2>&1; echo "pytest(PDL) rc=$?" >> gpurun_out/pytest_gpu_pdl.log
for pdl in 0 1; do
  B200COLL_PDL=$pdl timeout 60 ./build/b200coll_perf --devs 0,0 --op all_reduce -b 1K -e 1M -f 4 --iters 50 --warmup 10 > gpurun_out/pdl${pdl}_v2.txt 2>&1
done
tail -2 gpurun_out/pytest_gpu_final.log; tail -2 gpurun_out/pytest_gpu_pdl.log; paste <(grep -v "^#" gpurun_out/pdl0_v2.txt | awk '{print $1,$4,$6}') <(grep -v "^#" gpurun_out/pdl1_v2.txt | awk '{print $6}')
