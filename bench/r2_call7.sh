#!/bin/bash
# Round 2, calls 7 / 8 (4 or 8 GPUs): the final tree — bench.py both arms with the driver's step counts, the re-shaped rooted NVLS ops.
NG=${1:-4}; BUDGET_S=${2:-150}
mkdir -p gpurun_out; O=gpurun_out/r2c7_n${NG}
export B200COLL_TIMEOUT_MS=8000
T0=$SECONDS; left() { echo $((BUDGET_S - (SECONDS - T0))); }; ok() { [ $(left) -gt ${1:-30} ]; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
ALL=$(python3 -c "print(','.join(str(i) for i in range($NG)))")
port() { echo $((29500 + RANDOM % 1000)); }
timeout 100 $TR --master-port $(port) bench.py --gpus $NG --steps 20 --warmup 5 > ${O}_bench_ours.json 2> ${O}_bench_ours.err; echo "bench ours rc=$? ($(left) s left)"
timeout 100 $TR --master-port $(port) bench.py --gpus $NG --steps 20 --warmup 5 --impl reference > ${O}_bench_reference.json 2> ${O}_bench_reference.err; echo "bench reference rc=$? ($(left) s left)"
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
    print("ratio value", round(o["value"] / r["value"], 3), "e2e", round(o["e2e"]["value"] / r["e2e"]["value"], 3))
    for a, b, c, d in zip(o["table"], r["table"], o["e2e"]["table"], r["e2e"]["table"]):
        print(a["bytes"], a["algo"], a["oop_us"], b["oop_us"], "| e2e", c["e2e_us"], d["e2e_us"])
PY
grep -c "NCCL INFO" ${O}_bench_reference.err | sed 's/^/NCCL INFO lines on stderr: /'; grep -m 2 -E "Init COMPLETE|nranks" ${O}_bench_reference.err | cut -c 1-200
for op in broadcast reduce; do
  for shape in default "B200COLL_ROOTED_CTAS=64" "B200COLL_ROOTED_CTAS=296"; do
    ok 25 || break
    tag=$(echo "$shape" | tr -c 'A-Za-z0-9=' '_')
    env $( [ "$shape" = default ] || echo $shape ) timeout 40 ./build/b200coll_perf --devs $ALL --procs --op $op -b 1M -e 1G -f 8 --iters 10 --warmup 3 -c 1 > ${O}_rooted_${op}_$tag.txt 2>&1
    echo "$op [$shape] rc=$?: $(grep -E '^ +[0-9]' ${O}_rooted_${op}_$tag.txt | awk '{printf "%s:%s/%s ", $1, $6, $8}') $(grep -E 'Out of bounds' ${O}_rooted_${op}_$tag.txt)"
  done
done
if ok 40; then
  timeout 60 python -m pytest tests/test_coll_gpu.py -q -k "multi_gpu" --timeout 120 -p no:cacheprovider > ${O}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -n 2 ${O}_pytest_multi.log
fi
echo "done, $(left) s left"
