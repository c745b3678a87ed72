#!/bin/bash
# Round 2, last call (1 GPU): regression pass over the final tree.
mkdir -p gpurun_out; O=gpurun_out/r2c8
export B200COLL_TIMEOUT_MS=8000
ls /dev/nvidia* > ${O}_dev.txt 2>&1
timeout 150 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider -rs > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 14 ${O}_pytest.log | cut -c 1-300
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -n 1 ${O}_smoke.txt
