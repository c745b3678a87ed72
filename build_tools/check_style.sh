#!/bin/bash
# Presubmit style gate (role of reference build/check_gofmt.sh + build/check_errorf.sh, Makefile:27-35):
#   * every Python file byte-compiles, every shell script passes `bash -n`
#   * no tabs / trailing whitespace in Python sources (the gofmt analogue for this tree)
#   * error strings handed to set_last_error()/raise start lower-case unless they quote a reference message verbatim
#     (the reference's errorf rule: forbid `fmt.Errorf("Xxx`); an allow-list covers the mirrored kubelet-facing strings)
set -u
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
rc=0
while IFS= read -r f; do
  python -m py_compile "$f" 2>/dev/null || { echo "py_compile failed: ${f#$ROOT/}"; rc=1; }
  if grep -nP '\t| +$' "$f" >/dev/null; then echo "tab or trailing whitespace: ${f#$ROOT/}"; rc=1; fi
done < <(find "$ROOT" -name '*.py' -not -path '*/gpurun_out/*' -not -path '*/baseline/*' -not -path '*/build/*')
while IFS= read -r f; do
  bash -n "$f" || { echo "bash -n failed: ${f#$ROOT/}"; rc=1; }
done < <(find "$ROOT" -name '*.sh' -not -path '*/gpurun_out/*' -not -path '*/baseline/*')
ALLOW='Number of partitions|Not all GPUs|MaxSharedClientsPerGPU|GPU sharing strategy|Invalid HealthCriticalXid|NVIDIA MPS|GPU device|Compute instance|NVLS required'
if grep -rnE 'set_last_error\("[A-Z][a-z]' "$ROOT/coll/src" | grep -vE "$ALLOW"; then echo "capitalised error string in coll/src (see above)"; rc=1; fi
exit $rc
