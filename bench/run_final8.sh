#!/bin/bash
# Final 8-GPU record: all four collectives, both arms in one process group each, NCCL defaults for the headline op, and the
# expert-dispatch all-to-all-v benchmark.
NG=${1:-8}
mkdir -p gpurun_out; export B200COLL_TIMEOUT_MS=5000
O=gpurun_out/f${NG}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
OPS=all_gather,reduce_scatter,alltoall
timeout 240 $TR --master-port 29721 bench.py --gpus $NG --steps 20 --warmup 5 --table --extra-ops $OPS --extra-out ${O}_extra_ours.json > ${O}_bench.json 2> ${O}_bench.err
timeout 240 $TR --master-port 29722 bench.py --gpus $NG --steps 20 --warmup 5 --table --no-e2e --impl reference --extra-ops $OPS --extra-out ${O}_extra_ref.json > ${O}_ref.json 2> ${O}_ref.err
B200_REF_PROFILE=0 timeout 200 $TR --master-port 29723 bench.py --gpus $NG --steps 20 --warmup 5 --table --no-e2e --impl reference > ${O}_ref_defaults.json 2> ${O}_ref_defaults.err
timeout 200 $TR --master-port 29724 bench/alltoallv_perf.py > ${O}_alltoallv.jsonl 2> ${O}_alltoallv.err
ALL=$(python3 -c "print(','.join(str(i) for i in range($NG)))")
timeout 90 ./build/b200coll_perf --devs $ALL --procs --op all_reduce -b 1K -e 256M -f 2 --iters 10 --warmup 3 --json ${O}_perf.jsonl > ${O}_perf_ar.txt 2>&1; tail -1 ${O}_perf_ar.txt
grep -h "Avg bus" ${O}_*.err; cat ${O}_alltoallv.jsonl | cut -c1-400
