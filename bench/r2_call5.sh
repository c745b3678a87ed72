#!/bin/bash
# Round 2, call 5 (8 GPUs, charged 8x): the numbers that go into profiles/. Ordered by importance; every step has its own timeout and the
# script stops starting new sections once BUDGET_S seconds have passed.
NG=${1:-8}; BUDGET_S=${2:-560}
mkdir -p gpurun_out; O=gpurun_out/r2c5_n${NG}
export B200COLL_TIMEOUT_MS=8000
T0=$SECONDS
left() { echo $((BUDGET_S - (SECONDS - T0))); }
ok() { [ $(left) -gt ${1:-30} ]; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
ALL=$(python3 -c "print(','.join(str(i) for i in range($NG)))")
port() { echo $((29500 + RANDOM % 1000)); }
echo "== $(date -u +%T) bench.py ours / reference"
timeout 150 $TR --master-port $(port) bench.py --gpus $NG --steps 10 --warmup 3 > ${O}_bench_ours.json 2> ${O}_bench_ours.err; echo "bench ours rc=$? ($(left) s left)"
timeout 150 $TR --master-port $(port) bench.py --gpus $NG --steps 10 --warmup 3 --impl reference > ${O}_bench_reference.json 2> ${O}_bench_reference.err; echo "bench reference rc=$? ($(left) s left)"
python - $O <<'PY'
import json, sys
O = sys.argv[1]
res = {}
for arm in ("ours", "reference"):
    try:
        d = json.loads([l for l in open(f"{O}_bench_{arm}.json").read().splitlines() if l.startswith("{")][-1]); res[arm] = d
        print(arm, "value", d["value"], "peak", d["peak_busbw"], "e2e", d["e2e"]["value"], "verified", d["verified_vs_torch_fp32"], d.get("backend"))
    except Exception as e:
        print(arm, "no result", e)
if len(res) == 2:
    o, r = res["ours"], res["reference"]
    for a, b, c, d in zip(o["table"], r["table"], o["e2e"]["table"], r["e2e"]["table"]):
        print(a["bytes"], a["algo"], a["oop_us"], b["oop_us"], "| e2e", c["e2e_us"], d["e2e_us"])
PY
if ok 60; then
echo "== $(date -u +%T) host path per size"
timeout 90 $TR --master-port $(port) bench/e2e_hostpath.py --legs --iters 3 --tag n$NG > ${O}_host.jsonl 2> ${O}_host.err; cat ${O}_host.jsonl
fi
if ok 60; then
echo "== $(date -u +%T) A/B all_reduce: multicast barrier, PDL"
for cfg in "default" "B200COLL_MCBAR=0" "B200COLL_PDL=0"; do
  ok 25 || break
  tag=$(echo "$cfg" | tr -c 'A-Za-z0-9=' '_')
  env $( [ "$cfg" = default ] || echo $cfg ) timeout 60 ./build/b200coll_perf --devs $ALL --procs --op all_reduce -b 1K -e 64M -f 4 --iters 100 --warmup 10 -c 0 > ${O}_ab_ar_$tag.txt 2>&1
  echo "all_reduce [$cfg] rc=$?: $(grep -E '^ +[0-9]' ${O}_ab_ar_$tag.txt | awk '{printf "%s:%s ", $1, $6}')"
done
fi
if ok 60; then
echo "== $(date -u +%T) A/B copy engine"
for op in all_gather alltoall; do
  for cfg in "default" "B200COLL_BULK=0"; do
    ok 25 || break
    tag=$(echo "$cfg" | tr -c 'A-Za-z0-9=' '_')
    env $( [ "$cfg" = default ] || echo $cfg ) timeout 60 ./build/b200coll_perf --devs $ALL --procs --op $op -b 4M -e 1G -f 4 --iters 10 --warmup 3 -c 1 > ${O}_ab_${op}_$tag.txt 2>&1
    echo "$op [$cfg] rc=$?: $(grep -E '^ +[0-9]' ${O}_ab_${op}_$tag.txt | awk '{printf "%s:%s/%s ", $1, $6, $8}') $(grep -E 'Out of bounds' ${O}_ab_${op}_$tag.txt)"
  done
done
fi
if ok 100; then
echo "== $(date -u +%T) every other collective, both arms"
for impl in ours reference; do
  ok 60 || break
  timeout 110 $TR --master-port $(port) bench.py --gpus $NG --steps 10 --warmup 3 --no-e2e --impl $impl --op all_gather --min 64K --extra-ops reduce_scatter,alltoall,broadcast,reduce,sendrecv,gather,scatter --extra-out ${O}_ops_$impl.json > ${O}_ops_ag_$impl.json 2> ${O}_ops_$impl.err; echo "ops $impl rc=$? ($(left) s left)"
done
python - $O <<'PY'
import json, sys
O = sys.argv[1]
for impl in ("ours", "reference"):
    try:
        d = json.loads([l for l in open(f"{O}_ops_ag_{impl}.json").read().splitlines() if l.startswith("{")][-1])
        print(impl, "all_gather avg", d["value"], "peak", d["peak_busbw"], "verified", d["verified_vs_torch_fp32"])
        for op, v in json.load(open(f"{O}_ops_{impl}.json"))["ops"].items():
            print(impl, op, "avg", v["avg_busbw"], "peak", v["peak_busbw"], "verified", v["verified"])
    except Exception as e:
        print(impl, "no result", e)
PY
fi
if ok 50; then
echo "== $(date -u +%T) bench.py reference-sym (no e2e)"
timeout 100 $TR --master-port $(port) bench.py --gpus $NG --steps 10 --warmup 3 --no-e2e --impl reference-sym > ${O}_bench_reference-sym.json 2> ${O}_bench_reference-sym.err; echo "bench reference-sym rc=$? ($(left) s left)"
fi
if ok 60; then
echo "== $(date -u +%T) one ncu per rank (single-pass set)"
M2="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_ltcfabric.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size"
for spec in "ar_nvls all_reduce auto 64M k_ar_nvls" "ag all_gather auto 64M k_bulk|k_ag" "rs reduce_scatter auto 64M k_pull" "ll all_reduce auto 1K k_ll"; do
  ok 40 || break
  set -- $spec
  NCU_TIMEOUT=45 NCU_SKIP=3 NCU_COUNT=1 bash bench/ncu_ranks.sh $NG n${NG}_$1_mem "$5" "$M2" ./build/b200coll_perf --op $2 --algo $3 -b $4 -e $4 --iters 3 --warmup 3 -c 0 > /dev/null
  echo "ncu $1 done ($(left) s left)"
done
fi
if ok 45; then
echo "== $(date -u +%T) host path without NUMA placement"
B200COLL_AFFINITY=0 timeout 60 $TR --master-port $(port) bench/e2e_hostpath.py --min $((256<<20)) --iters 3 --torch-pinned --tag n${NG}_noaffinity_torchpinned > ${O}_host_noaff.jsonl 2> ${O}_host_noaff.err; cat ${O}_host_noaff.jsonl
fi
echo "== $(date -u +%T) done, $(left) s of budget left"
