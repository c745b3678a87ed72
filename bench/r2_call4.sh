#!/bin/bash
# Round 2, call 4 (1 GPU): new tests on hardware, smoke() plain and under a profiler, sanitizer x copy-engine kernel, ncu probes,
# bench both arms, one full ncu capture of the one-rank copy kernel, and (last) the MIG attempt.
mkdir -p gpurun_out; O=gpurun_out/r2c4
export B200COLL_TIMEOUT_MS=8000
echo "== $(date -u +%T) pytest"
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 30 ${O}_pytest.log | cut -c 1-300
echo "== $(date -u +%T) smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -n 3 ${O}_smoke.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${O}_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke_ncu.txt 2>&1; echo "smoke under ncu rc=$?"; tail -n 2 ${O}_smoke_ncu.txt
python - <<'PY'
import csv, collections
c = collections.Counter()
try:
    for r in csv.DictReader(l for l in open("gpurun_out/r2c4_smoke_launches.csv") if l.startswith('"')):
        c[r["Kernel Name"].split("<")[0].split("(")[0].replace("void ", "")] += 1
    print("kernel families under ncu:", dict(c))
except Exception as e:
    print("no launch list:", e)
PY
echo "== $(date -u +%T) sanitizer x k_bulk"
for cfg in "1rank --devs 0 --op all_reduce -b 2M -e 2M" "2virt_ag --devs 0,0 --op all_gather -b 4M -e 4M" "2virt_a2a --devs 0,0 --op alltoall -b 4M -e 4M"; do
  set -- $cfg; tag=$1; shift
  for bulk in 1 0; do
    B200COLL_BULK=$bulk B200COLL_TIMEOUT_MS=40000 timeout 90 compute-sanitizer --tool memcheck --print-limit 5 ./build/b200coll_perf "$@" --iters 1 --warmup 1 -c 1 > ${O}_san_${tag}_bulk$bulk.log 2>&1
    echo "memcheck $tag bulk=$bulk rc=$? $(grep -E 'ERROR SUMMARY|WATCHDOG|Out of bounds' ${O}_san_${tag}_bulk$bulk.log | tr '\n' ' ')"
  done
done
B200COLL_TIMEOUT_MS=40000 timeout 90 compute-sanitizer --tool racecheck --print-limit 5 ./build/b200coll_perf --devs 0 --op all_reduce -b 2M -e 2M --iters 1 --warmup 1 -c 1 > ${O}_san_race_1rank.log 2>&1; echo "racecheck 1rank bulk rc=$? $(grep -E 'RACECHECK SUMMARY' ${O}_san_race_1rank.log)"
echo "== $(date -u +%T) ncu probes"
timeout 120 ncu --metrics nvlrx__bytes.sum,nvltx__bytes.sum --clock-control none --cache-control none -k regex:k_bulk -c 1 ./build/b200coll_perf --devs 0 --op all_reduce -b 64M -e 64M --iters 2 --warmup 1 -c 0 > ${O}_ncu_nvl_probe.txt 2>&1; echo "nvlink metrics on one GPU rc=$?"; grep -E "nvl|ERROR" ${O}_ncu_nvl_probe.txt | head -5
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_bulk -s 2 -c 1 -o gpurun_out/prof_bulk_copy ./build/b200coll_perf --devs 0 --op all_reduce -b 1G -e 1G --iters 3 --warmup 2 -c 0 > ${O}_ncu_full_bulk.txt 2>&1; echo "ncu --set full k_bulk rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_copy_scale -s 2 -c 1 -o gpurun_out/prof_copy_scale_cast ./build/b200coll_perf --devs 0 --op all_reduce --out-dtype float -b 512M -e 512M --iters 3 --warmup 2 -c 0 > ${O}_ncu_full_cast.txt 2>&1; echo "ncu --set full k_copy_scale(bf16->f32) rc=$?"
echo "== $(date -u +%T) bench both arms"
timeout 300 python bench.py --steps 10 --warmup 3 > ${O}_bench_ours.json 2> ${O}_bench_ours.err; echo "bench ours rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --impl reference > ${O}_bench_ref.json 2> ${O}_bench_ref.err; echo "bench ref rc=$?"
B200COLL_BULK=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e > ${O}_bench_ours_nobulk.json 2> ${O}_bench_ours_nobulk.err; echo "bench ours (BULK=0) rc=$?"
python - <<'PY'
import json
for arm in ("ours", "ref", "ours_nobulk"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2c4_bench_{arm}.json").read().splitlines() if l.startswith("{")][-1])
        big = {r["bytes"]: r["oop_busbw"] for r in d["table"]}
        print(arm, "value", d["value"], "peak", d["peak_busbw"], "e2e", (d.get("e2e") or {}).get("value"), "verified", d["verified_vs_torch_fp32"], "| algbw at 16M/256M/1G:", big.get(1 << 24), big.get(1 << 28), big.get(1 << 30))
    except Exception as e:
        print(arm, "no result", e)
PY
B200COLL_HOST_ZEROCOPY_KB=65536 timeout 120 python bench/e2e_hostpath.py --min $((256<<10)) --max $((64<<20)) --factor 2 --tag zerocopy_up_to_64M > ${O}_host_zc64m.jsonl 2>&1; grep bytes ${O}_host_zc64m.jsonl | python -c "import sys,json; [print(d['bytes'], d['seq_us'], d['lib_us']) for d in map(json.loads, sys.stdin)]"
echo "== $(date -u +%T) MIG attempt (last)"
timeout 400 bash bench/r2_mig_attempt.sh
nvidia-smi --query-gpu=index,mig.mode.current,mig.mode.pending --format=csv
echo "== $(date -u +%T) done"
