#!/bin/bash
# Round 2, call 3 (N GPUs, default 2): multi-GPU tests, bench both arms, A/B of the latency and copy-engine switches, sanitizers on
# real peers, one ncu per rank. Sections: tests bench ab san ncu e2e
NG=${1:-2}; SECTIONS=",${2:-tests,bench,ab,san,ncu,e2e},"
want() { case "$SECTIONS" in *",$1,"*) return 0;; *) return 1;; esac; }
mkdir -p gpurun_out; O=gpurun_out/r2c3_n${NG}
export B200COLL_TIMEOUT_MS=8000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
ALL=$(python3 -c "print(','.join(str(i) for i in range($NG)))")
if want tests; then
echo "== $(date -u +%T) pytest -m gpu on $NG GPUs"
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 25 ${O}_pytest.log
fi
if want bench; then
echo "== $(date -u +%T) bench.py three arms"
for impl in ours reference reference-sym; do
  timeout 400 $TR --master-port $((29500 + RANDOM % 400)) bench.py --gpus $NG --steps 10 --warmup 3 --impl $impl > ${O}_bench_$impl.json 2> ${O}_bench_$impl.err; echo "bench $impl rc=$?"
done
python - $O <<'PY'
import json, sys
O = sys.argv[1]
res = {}
for arm in ("ours", "reference", "reference-sym"):
    try:
        d = json.loads([l for l in open(f"{O}_bench_{arm}.json").read().splitlines() if l.startswith("{")][-1]); res[arm] = d
        print(arm, "value", d["value"], "peak", d["peak_busbw"], "e2e", d["e2e"]["value"], "verified", d["verified_vs_torch_fp32"], d.get("backend"))
    except Exception as e:
        print(arm, "no result", e)
if "ours" in res and "reference" in res:
    o, r = res["ours"], res["reference"]
    print("bytes | ours us | ref us | ours e2e us | ref e2e us")
    for a, b, c, d in zip(o["table"], r["table"], o["e2e"]["table"], r["e2e"]["table"]):
        print(a["bytes"], a["oop_us"], b["oop_us"], c["e2e_us"], d["e2e_us"])
PY
fi
if want ab; then
echo "== $(date -u +%T) A/B: multicast barrier, PDL, copy engine"
for cfg in "default" "B200COLL_MCBAR=0" "B200COLL_PDL=0" "B200COLL_MCBAR=0 B200COLL_PDL=0"; do
  tag=$(echo "$cfg" | tr -c 'A-Za-z0-9=' '_')
  env $( [ "$cfg" = default ] || echo $cfg ) timeout 120 ./build/b200coll_perf --devs $ALL --procs --op all_reduce -b 1K -e 16M -f 4 --iters 200 --warmup 20 -c 0 > ${O}_ab_ar_$tag.txt 2>&1
  echo "all_reduce [$cfg] rc=$?: $(grep -E '^ +[0-9]' ${O}_ab_ar_$tag.txt | awk '{printf "%s:%s ", $1, $4}')"
done
for op in all_gather alltoall broadcast; do
  for cfg in "default" "B200COLL_BULK=0"; do
    tag=$(echo "$cfg" | tr -c 'A-Za-z0-9=' '_')
    env $( [ "$cfg" = default ] || echo $cfg ) timeout 120 ./build/b200coll_perf --devs $ALL --procs --op $op -b 1M -e 1G -f 4 --iters 20 --warmup 5 -c 1 > ${O}_ab_${op}_$tag.txt 2>&1
    echo "$op [$cfg] rc=$?: $(grep -E '^ +[0-9]' ${O}_ab_${op}_$tag.txt | awk '{printf "%s:%s/%s ", $1, $4, $6}') $(grep -E 'errors|wrong|Out of bounds' ${O}_ab_${op}_$tag.txt | tail -n 1)"
  done
done
for cap in 148 296 444; do
  B200COLL_BULK_CTAS=$cap timeout 120 ./build/b200coll_perf --devs $ALL --procs --op all_gather -b 64M -e 1G -f 4 --iters 20 --warmup 5 -c 0 > ${O}_ab_agctas_$cap.txt 2>&1
  echo "all_gather bulk ctas=$cap: $(grep -E '^ +[0-9]' ${O}_ab_agctas_$cap.txt | awk '{printf "%s:%s/%s ", $1, $4, $6}')"
done
fi
if want san; then
echo "== $(date -u +%T) compute-sanitizer on real peers (racecheck, memcheck, synccheck)"
export B200COLL_TIMEOUT_MS=180000
: > ${O}_sanitizer.log
for tool in racecheck memcheck synccheck; do
  for spec in "all_reduce auto" "all_reduce ll" "all_reduce ll2" "all_reduce oneshot" "all_reduce twoshot" "all_reduce nvls" "all_gather auto" "reduce_scatter auto" "alltoall auto" "broadcast auto" "reduce auto" "sendrecv auto"; do
    set -- $spec
    [ $tool != racecheck ] && [ "$2" != auto ] && [ "$2" != nvls ] && continue
    echo "--- $tool $1 $2" >> ${O}_sanitizer.log
    timeout 150 compute-sanitizer --tool $tool --print-limit 5 ./build/b200coll_perf --devs $ALL --op $1 --algo $2 -b 4K -e 1M -f 16 --iters 1 --warmup 1 -c 1 >> ${O}_sanitizer.log 2>&1
    echo "rc=$?" >> ${O}_sanitizer.log
  done
done
grep -E "^--- |ERROR SUMMARY|RACECHECK SUMMARY|rc=" ${O}_sanitizer.log | paste - - - | head -60
export B200COLL_TIMEOUT_MS=8000
fi
if want ncu; then
echo "== $(date -u +%T) one ncu per rank"
M1="nvlrx__bytes.sum,nvltx__bytes.sum,gpu__time_duration.sum"
M2="dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_ltcfabric.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size"
M3="smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_membar.ratio,smsp__average_warp_latency_issue_stalled_barrier.ratio,smsp__average_warp_latency_issue_stalled_lg_throttle.ratio,smsp__inst_executed.sum"
for spec in "ar_auto all_reduce auto 64M k_ar" "ar_nvls all_reduce nvls 64M k_ar_nvls" "ag all_gather auto 64M k_bulk|k_ag" "a2a alltoall auto 64M k_bulk|k_a2av" "rs reduce_scatter auto 64M k_pull" "ll all_reduce auto 1K k_ll" "bcast broadcast auto 64M k_bcast|k_bulk"; do
  set -- $spec
  NCU_SKIP=3 NCU_COUNT=2 bash bench/ncu_ranks.sh $NG n${NG}_$1_nvl "$5" "$M1" ./build/b200coll_perf --op $2 --algo $3 -b $4 -e $4 --iters 3 --warmup 3 -c 0 > /dev/null
  NCU_SKIP=3 NCU_COUNT=2 bash bench/ncu_ranks.sh $NG n${NG}_$1_mem "$5" "$M2" ./build/b200coll_perf --op $2 --algo $3 -b $4 -e $4 --iters 3 --warmup 3 -c 0 > /dev/null
  NCU_SKIP=3 NCU_COUNT=2 bash bench/ncu_ranks.sh $NG n${NG}_$1_stall "$5" "$M3" ./build/b200coll_perf --op $2 --algo $3 -b $4 -e $4 --iters 3 --warmup 3 -c 0 > /dev/null
  echo "ncu $1: $(tail -n 2 gpurun_out/ncu_n${NG}_$1_nvl_r0.csv | cut -c 1-400)"
done
fi
if want e2e; then
echo "== $(date -u +%T) host path per size"
timeout 300 $TR --master-port $((29500 + RANDOM % 400)) bench/e2e_hostpath.py --legs --tag n$NG > ${O}_host.jsonl 2> ${O}_host.err; cat ${O}_host.jsonl
B200COLL_AFFINITY=0 timeout 300 $TR --master-port $((29500 + RANDOM % 400)) bench/e2e_hostpath.py --min $((64<<20)) --torch-pinned --tag n${NG}_noaffinity_torchpinned > ${O}_host_noaff.jsonl 2> ${O}_host_noaff.err; cat ${O}_host_noaff.jsonl
fi
echo "== $(date -u +%T) done"
