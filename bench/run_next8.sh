#!/bin/bash
# Everything that was written after the GPU time of round 1 ran out, or needs more than one GPU. Sections (second argument, comma
# separated; default = all):
#   tests1    virtual-rank tests of send/recv and CommSplit, the process group's CUDA paths, pipelined e2e, tools   (any box, 1 GPU is enough)
#   multi     multi-process NVLS tests a one-GPU box skips                                                         (>= 2 GPUs)
#   rooted    broadcast / reduce on the NVLS path, both arms                                                       (>= 3 GPUs for multicast)
#   pdl       B200COLL_PDL=0/1 small-message latency                                                               (>= 2 GPUs)
#   variants  shipped build vs libb200coll_{gridconst,mcbar,bulk}.so, all_reduce + all_gather                      (mcbar needs multicast)
#   e2e       copy-then-reduce vs Comm.all_reduce_from_host                                                        (>= 2 GPUs)
#   p2p       sendrecv on both arms + CTAs-per-operation sweep                                                     (>= 2 GPUs)
#   ddp       DDP demo: arena pool / plain / NCCL                                                                  (>= 2 GPUs)
# GPU-minutes are charged per GPU, so spend them in this order:
#   gpurun --timeout 900 -- 'bash bench/run_next8.sh 1 tests1'                                  # ~6 min x 1
#   gpurun --gpus 2 --timeout 900 -- 'bash bench/run_next8.sh 2 multi,pdl,e2e,p2p,ddp'          # ~8 min x 2
#   gpurun --gpus 8 --timeout 600 -- 'bash bench/run_next8.sh 8 multi,rooted,pdl,variants,p2p'  # ~8 min x 8: the numbers that go into profiles/
NG=${1:-8}
mkdir -p gpurun_out; export B200COLL_TIMEOUT_MS=5000
O=gpurun_out/x${NG}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
ALL=$(python3 -c "print(','.join(str(i) for i in range($NG)))")
SECTIONS=",${2:-tests1,multi,rooted,pdl,variants,e2e,p2p,ddp},"
want() { case "$SECTIONS" in *",$1,"*) return 0;; *) return 1;; esac; }
if want multi; then
echo "== $(date -u +%T) multi-process NVLS tests"
timeout 300 python -m pytest tests/test_coll_gpu.py -q -k "multi_gpu" > ${O}_pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -n 2 ${O}_pytest_multi.log
fi
if want tests1; then
echo "== $(date -u +%T) torch.distributed backend on CUDA (first run on hardware)"
B200_RUN_UNVALIDATED=1 timeout 300 python -m pytest tests/test_process_group.py -q -m gpu > ${O}_pytest_pg.log 2>&1; echo "pytest rc=$?"; tail -n 3 ${O}_pytest_pg.log
fi
if want rooted; then
echo "== $(date -u +%T) rooted ops, both arms"
timeout 200 $TR --master-port 29731 bench.py --gpus $NG --op broadcast --steps 20 --warmup 5 --table --no-e2e --extra-ops reduce --extra-out ${O}_rooted_ours.json > ${O}_bcast.json 2> ${O}_bcast.err
timeout 200 $TR --master-port 29732 bench.py --gpus $NG --op broadcast --steps 20 --warmup 5 --table --no-e2e --impl reference --extra-ops reduce --extra-out ${O}_rooted_ref.json > ${O}_bcast_ref.json 2> ${O}_bcast_ref.err
grep -h "Avg bus" ${O}_bcast.err ${O}_bcast_ref.err
fi
if want pdl; then
echo "== $(date -u +%T) PDL off / on, all_reduce 1 KiB .. 4 MiB"
for pdl in 0 1; do
  B200COLL_PDL=$pdl timeout 90 ./build/b200coll_perf --devs $ALL --procs --op all_reduce -b 1K -e 4M -f 4 --iters 200 --warmup 20 > ${O}_pdl${pdl}.txt 2>&1; echo "pdl=$pdl rc=$?"; grep -E "^ +[0-9]" ${O}_pdl${pdl}.txt | awk '{print $1, $4, $6}' | tr '\n' ';'; echo
done
fi
if want variants; then
echo "== $(date -u +%T) build variants (A/B candidates: kernel parameter in the constant bank; multicast barrier)"
if [ -f coll/lib/libb200coll_gridconst.so ] && [ -f coll/lib/libb200coll_mcbar.so ] && [ -f coll/lib/libb200coll_bulk.so ]; then echo "variants: prebuilt libraries travelled with the snapshot"
else make -C coll variants -j2 > ${O}_variants_build.log 2>&1; echo "variants rc=$?"; fi
port=29740
for v in "" _gridconst _mcbar _bulk; do
  lib=coll/lib/libb200coll${v}.so
  [ -f $lib ] || continue
  port=$((port + 1))
  B200COLL_LIB=$PWD/$lib timeout 200 $TR --master-port $port bench.py --gpus $NG --steps 20 --warmup 5 --no-e2e --max 1G --extra-ops all_gather --extra-out ${O}_ab${v:-_shipped}_ag.json > ${O}_ab${v:-_shipped}.json 2> ${O}_ab${v:-_shipped}.err
  python3 - "$lib" ${O}_ab${v:-_shipped}.json ${O}_ab${v:-_shipped}_ag.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    t = {r["bytes"]: r["oop_us"] for r in d["table"]}
    print(sys.argv[1], "all_reduce avg busbw", d["value"], "peak", d["peak_busbw"], "verified", d["verified_vs_torch_fp32"], "| us at 1K/64K/1M/16M:", t.get(1024), t.get(65536), t.get(1 << 20), t.get(1 << 24))
    ag = json.load(open(sys.argv[3]))["ops"]["all_gather"]
    big = {r["bytes"]: r["oop_busbw"] for r in ag["table"]}
    print(sys.argv[1], "all_gather avg busbw", round(ag["avg_busbw"], 1), "peak", round(ag["peak_busbw"], 1), "verified", ag["verified"], "| busbw at 16M/256M/1G:", big.get(1 << 24), big.get(1 << 28), big.get(1 << 30))
except Exception as e:
    print(sys.argv[1], "no result:", e)
PY
done
fi
if want tests1; then
echo "== $(date -u +%T) virtual-rank tests: send/recv, CommSplit, pipelined host all-reduce"
B200_RUN_UNVALIDATED=1 timeout 600 python -m pytest tests/test_coll_gpu.py -q -k "send_recv or sendrecv or comm_split or all_reduce_from_host" > ${O}_pytest_p2p.log 2>&1; echo "pytest rc=$?"; tail -n 3 ${O}_pytest_p2p.log
fi
if want e2e; then
echo "== $(date -u +%T) end-to-end step: copy-then-reduce vs pipelined Comm.all_reduce_from_host"
timeout 200 $TR --master-port 29760 bench/e2e_pipeline.py > ${O}_e2e_pipeline.jsonl 2> ${O}_e2e_pipeline.err; cat ${O}_e2e_pipeline.jsonl
fi
if want p2p; then
echo "== $(date -u +%T) sendrecv on both arms"
for impl in reference ours; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) bench.py --gpus $NG --steps 20 --warmup 5 --no-e2e --op sendrecv --extra-ops gather,scatter --extra-out ${O}_p2p_extra_$impl.json --impl $impl > ${O}_sendrecv_$impl.json 2> ${O}_sendrecv_$impl.err; echo "sendrecv $impl rc=$?"
done
echo "--- CTAs per send/recv operation (B200COLL_P2P_MAX_BLOCKS): pick the default from this"
for cap in 8 16 32; do
  B200COLL_P2P_MAX_BLOCKS=$cap timeout 200 ./build/sendrecv_perf --devs $ALL --procs -b 64K -e 1G -f 4 -w 3 -n 10 -c 0 > ${O}_sendrecv_cap$cap.txt 2>&1; echo "cap=$cap rc=$?"; tail -n 4 ${O}_sendrecv_cap$cap.txt
done
fi
if want ddp; then
echo "== $(date -u +%T) DDP demo: our backend (arena pool, plain) vs NCCL on the same box"
timeout 300 $TR --master-port $((29900 + RANDOM % 90)) demo/gpu-training/ddp_b200coll.py --steps 30 --arena-pool 2> ${O}_ddp_pool.err | tail -n 1 | tee ${O}_ddp_pool.json
for be in b200coll nccl; do
  timeout 300 $TR --master-port $((29900 + RANDOM % 90)) demo/gpu-training/ddp_b200coll.py --steps 30 --backend $be 2> ${O}_ddp_$be.err | tail -n 1 | tee ${O}_ddp_$be.json
done
fi
if want tests1; then
echo "== $(date -u +%T) tools on hardware (fault injector last: it kills its own context on purpose)"
B200_RUN_FAULT_INJECTION=1 timeout 300 python -m pytest tests/test_zz_tools_gpu.py -q -m gpu > ${O}_pytest_tools.log 2>&1; echo "pytest rc=$?"; tail -n 3 ${O}_pytest_tools.log
fi
echo "== $(date -u +%T) done"
