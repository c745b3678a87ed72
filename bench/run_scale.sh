#!/bin/bash
# bench.py at NG GPUs, both arms (+ NCCL defaults), plus the mid-size algorithm crossover for the tuner table.
NG=${1:-2}
mkdir -p gpurun_out; export B200COLL_TIMEOUT_MS=5000
P=./build/b200coll_perf
ALL=$(python3 -c "print(','.join(str(i) for i in range($NG)))")
O=gpurun_out/s${NG}
for algo in ll ll2 twoshot nvls; do
  timeout 60 $P --devs $ALL --procs --op all_reduce --algo $algo -b 64K -e 4M --iters 20 --warmup 5 > ${O}_ar_${algo}.txt 2>&1
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29711 bench.py --gpus $NG --steps 20 --warmup 5 --table > ${O}_bench.json 2> ${O}_bench.err
timeout 200 $TR --master-port 29712 bench.py --gpus $NG --steps 20 --warmup 5 --table --no-e2e --impl reference > ${O}_ref.json 2> ${O}_ref.err
B200_REF_PROFILE=0 timeout 200 $TR --master-port 29713 bench.py --gpus $NG --steps 20 --warmup 5 --table --no-e2e --impl reference > ${O}_ref_defaults.json 2> ${O}_ref_defaults.err
grep -h "Avg bus" ${O}_*.err
