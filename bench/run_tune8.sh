#!/bin/bash
# 8-GPU launch-shape + crossover tuning, one process group per family (shapes swept in-process).
NG=${1:-8}
mkdir -p gpurun_out; export B200COLL_TIMEOUT_MS=5000
P=./build/b200coll_perf
ALL=$(python3 -c "print(','.join(str(i) for i in range($NG)))")
O=gpurun_out/t${NG}
T0=$(date +%s)
t() { local s=$(date +%s); timeout 60 "$@"; echo "# rc=$? took $(( $(date +%s) - s ))s (t+$(( $(date +%s) - T0 ))s)"; }
NV="0:8:256,0:16:256,0:24:256,0:32:256,0:48:256,0:64:256,0:16:512,0:32:512,0:64:512,0:32:128,0:64:128,0:96:128"
PP="1:32:256,1:64:256,1:128:256,1:296:256,1:32:512,1:64:512,1:128:512,1:296:512"
t $P --devs $ALL --procs --op all_reduce --algo nvls -b 1M -e 1G -f 4 --iters 10 --warmup 3 --check 0 --inplace 0 --sweep $NV > ${O}_ar_nvls_shape.txt 2>&1
t $P --devs $ALL --procs --op all_reduce --algo ll2 -b 64K -e 4M --iters 20 --warmup 5 > ${O}_ar_ll2.txt 2>&1
t $P --devs $ALL --procs --op all_reduce --algo ll -b 64K -e 512K --iters 20 --warmup 5 > ${O}_ar_ll.txt 2>&1
t $P --devs $ALL --procs --op all_reduce --algo nvls -b 64K -e 16M --iters 20 --warmup 5 > ${O}_ar_nvls.txt 2>&1
t $P --devs $ALL --procs --op all_reduce --algo twoshot -b 4M -e 1G -f 16 --iters 10 --warmup 3 --check 0 --inplace 0 --sweep $PP > ${O}_ar_p2p_shape.txt 2>&1
for op in all_gather reduce_scatter; do
  t $P --devs $ALL --procs --op $op --algo twoshot -b 16M -e 1G -f 8 --iters 10 --warmup 3 --check 0 --inplace 0 --sweep $PP > ${O}_${op}_p2p_shape.txt 2>&1
  t $P --devs $ALL --procs --op $op --algo nvls -b 16M -e 1G -f 8 --iters 10 --warmup 3 --check 0 --inplace 0 --sweep $NV > ${O}_${op}_nvls_shape.txt 2>&1
done
t $P --devs $ALL --procs --op alltoall --algo twoshot -b 16M -e 1G -f 8 --iters 10 --warmup 3 --check 0 --inplace 0 --sweep $PP > ${O}_a2a_p2p_shape.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
t $TR --master-port 29701 bench.py --gpus $NG --steps 20 --warmup 5 --table > ${O}_bench.json 2> ${O}_bench.err
B200_REF_PROFILE=0 t $TR --master-port 29702 bench.py --gpus $NG --steps 20 --warmup 5 --table --no-e2e --impl reference > ${O}_ref_defaults.json 2> ${O}_ref_defaults.err
grep -h "Avg bus" ${O}_*.err; grep -h "took" ${O}_*.txt | head -20
