#!/bin/bash
# 8-GPU exploration: forced-algorithm sweeps (crossovers), NVLS launch-shape tuning, other ops, NCCL reference.
NG=${1:-8}
mkdir -p gpurun_out; export B200COLL_TIMEOUT_MS=5000
P=./build/b200coll_perf
ALL=$(python3 -c "print(','.join(str(i) for i in range($NG)))")
O=gpurun_out/n${NG}
t() { timeout 100 "$@"; }
J="--json ${O}.jsonl"
t $P --devs $ALL --procs --op all_reduce --algo ll -b 1K -e 512K --iters 20 --warmup 5 $J > ${O}_ar_ll.txt 2>&1
t $P --devs $ALL --procs --op all_reduce --algo oneshot -b 1K -e 4M --iters 20 --warmup 5 $J > ${O}_ar_oneshot.txt 2>&1
t $P --devs $ALL --procs --op all_reduce --algo twoshot -b 64K -e 1G --iters 20 --warmup 5 $J > ${O}_ar_twoshot.txt 2>&1
t $P --devs $ALL --procs --op all_reduce --algo nvls -b 64K -e 1G --iters 20 --warmup 5 $J > ${O}_ar_nvls.txt 2>&1
for th in 256 512; do for ctas in 32 64 148 296; do
  echo "## threads=$th ctas=$ctas" >> ${O}_nvls_shape.txt
  B200COLL_FORCE_THREADS=$th t $P --devs $ALL --procs --op all_reduce --algo nvls -b 4M -e 1G -f 16 --iters 10 --warmup 3 --check 0 --inplace 0 --max-ctas $ctas >> ${O}_nvls_shape.txt 2>&1
done; done
for op in all_gather reduce_scatter; do
  t $P --devs $ALL --procs --op $op --algo ll -b 8K -e 2M --iters 20 --warmup 5 $J > ${O}_${op}_ll.txt 2>&1
  t $P --devs $ALL --procs --op $op --algo twoshot -b 64K -e 1G -f 4 --iters 20 --warmup 5 $J > ${O}_${op}_p2p.txt 2>&1
  t $P --devs $ALL --procs --op $op --algo nvls -b 64K -e 1G -f 4 --iters 20 --warmup 5 $J > ${O}_${op}_nvls.txt 2>&1
done
t $P --devs $ALL --procs --op alltoall --algo ll -b 8K -e 2M --iters 20 --warmup 5 $J > ${O}_a2a_ll.txt 2>&1
t $P --devs $ALL --procs --op alltoall --algo twoshot -b 64K -e 1G -f 4 --iters 20 --warmup 5 $J > ${O}_a2a_p2p.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
port=29600
for op in all_reduce all_gather reduce_scatter alltoall; do
  port=$((port+1))
  timeout 200 $TR --master-port $port bench.py --gpus $NG --steps 20 --warmup 5 --table --no-e2e --op $op --impl reference > ${O}_ref_${op}.json 2> ${O}_ref_${op}.err
done
port=$((port+1))
timeout 300 $TR --master-port $port bench.py --gpus $NG --steps 20 --warmup 5 --table > ${O}_bench.json 2> ${O}_bench.err
grep -h "Avg bus" ${O}_*.err; grep -c . ${O}.jsonl; tail -2 ${O}_ar_nvls.txt
