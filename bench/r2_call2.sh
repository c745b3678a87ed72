#!/bin/bash
# Round 2, call 2 (1 GPU): the full GPU suite with the gates removed, bench.py both arms, and the host-path knobs.
mkdir -p gpurun_out; O=gpurun_out/r2c2
export B200COLL_TIMEOUT_MS=5000
timeout 600 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 25 ${O}_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > ${O}_bench_ours.json 2> ${O}_bench_ours.err; echo "bench ours rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --impl reference > ${O}_bench_ref.json 2> ${O}_bench_ref.err; echo "bench ref rc=$?"
python - <<'PY'
import json
for arm in ("ours", "ref"):
    try:
        d = json.loads(open(f"gpurun_out/r2c2_bench_{arm}.json").read().strip().splitlines()[-1])
        print(arm, "value", d["value"], "e2e", d["e2e"]["value"], "verified", d["verified_vs_torch_fp32"], "e2e rows:", [(r["bytes"], r["e2e_us"]) for r in d["e2e"]["table"][::4]])
    except Exception as e:
        print(arm, "no result", e)
PY
timeout 200 python bench/e2e_hostpath.py --legs --tag default > ${O}_host_default.jsonl 2> ${O}_host_default.err; echo "hostpath rc=$?"; cat ${O}_host_default.jsonl
B200COLL_HOST_ZEROCOPY_KB=512 B200COLL_HOST_PIPELINE_KB=512 timeout 200 python bench/e2e_hostpath.py --max $((16<<20)) --factor 2 --min $((16<<10)) --tag zc512_pipe512 > ${O}_host_zc512.jsonl 2>&1; cat ${O}_host_zc512.jsonl
B200COLL_HOST_ZEROCOPY_KB=0 timeout 200 python bench/e2e_hostpath.py --max $((1<<20)) --factor 2 --tag zc0 > ${O}_host_zc0.jsonl 2>&1; cat ${O}_host_zc0.jsonl
for ck in 1024 2048 4096 16384; do
  B200COLL_HOST_CHUNK_KB=$ck timeout 200 python bench/e2e_hostpath.py --min $((16<<20)) --tag chunk${ck}k > ${O}_host_c${ck}.jsonl 2>&1; grep bytes ${O}_host_c${ck}.jsonl | tr '\n' ' '; echo
done
timeout 100 python bench/e2e_hostpath.py --min $((16<<20)) --torch-pinned --tag torchpinned > ${O}_host_torchpinned.jsonl 2>&1; grep bytes ${O}_host_torchpinned.jsonl | tr '\n' ' '; echo
