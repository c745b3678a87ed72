#!/bin/bash
# Round 2, call 6 (1 GPU): regression pass over the final tree.
mkdir -p gpurun_out; O=gpurun_out/r2c6
export B200COLL_TIMEOUT_MS=8000
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 12 ${O}_pytest.log | cut -c 1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -n 2 ${O}_smoke.txt
timeout 200 python bench.py --steps 5 --warmup 3 > ${O}_bench_ours.json 2> ${O}_bench_ours.err; echo "bench rc=$?"; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2c6_bench_ours.json') if l.startswith('{')][-1]); print('value', d['value'], 'e2e', d['e2e']['value'], 'verified', d['verified_vs_torch_fp32'], 'launches', d['gpu_launches'])"
