#!/bin/bash
# Correctness matrix for libb200coll on whatever GPUs are visible. Every command runs under its own
# timeout so a protocol bug costs seconds, not the box. Usage: run_matrix.sh <outdir> [quick|full]
OUT=${1:-gpurun_out}; MODE=${2:-quick}
mkdir -p "$OUT"; export B200COLL_TIMEOUT_MS=${B200COLL_TIMEOUT_MS:-3000}
P=./build/b200coll_perf
NG=$(nvidia-smi -L | wc -l)
LOG=$OUT/matrix.log; : > "$LOG"
pass=0; fail=0
run() { # name, args...
  local name=$1; shift
  echo "=== $name: $*" >> "$LOG"
  if timeout 60 $P "$@" >> "$LOG" 2>&1; then pass=$((pass+1)); echo "PASS $name" >> "$OUT/matrix.summary"; else rc=$?; fail=$((fail+1)); echo "FAIL($rc) $name" >> "$OUT/matrix.summary"; fi
}
: > "$OUT/matrix.summary"
$P --selfcheck > "$OUT/selfcheck.txt" 2>&1
# N=1
run n1_ar --devs 0 --op all_reduce -b 1K -e 16M -f 16 --iters 5 --warmup 1 --scale 0.5
# virtual ranks on GPU 0 (P2P protocols only)
for n in 2 4; do
  devs=$(python3 -c "print(','.join(['0']*$n))")
  for algo in ll ll2 oneshot twoshot; do
    run v${n}_ar_${algo} --devs $devs --op all_reduce --algo $algo -b 1K -e 256K -f 4 --iters 5 --warmup 2
  done
  run v${n}_ar_tail --devs $devs --op all_reduce --algo twoshot -b 1030 -e 70000 -f 3 --iters 3 --warmup 1
  run v${n}_ar_ll_tail --devs $devs --op all_reduce --algo ll -b 1030 -e 70000 -f 3 --iters 3 --warmup 1
  run v${n}_ar_ll2_tail --devs $devs --op all_reduce --algo ll2 -b 1030 -e 700000 -f 3 --iters 3 --warmup 1
  for op in all_gather reduce_scatter alltoall; do
    run v${n}_${op}_ll --devs $devs --op $op --algo ll -b 1K -e 256K -f 4 --iters 5 --warmup 2
    run v${n}_${op}_p2p --devs $devs --op $op --algo twoshot -b 1K -e 4M -f 8 --iters 5 --warmup 2
  done
  run v${n}_ar_f32_bf16 --devs $devs --op all_reduce --algo twoshot --dtype f32 --out-dtype bf16 --scale 0.5 -b 4K -e 1M -f 16 --iters 3 --warmup 1
  run v${n}_ar_bf16_fp8 --devs $devs --op all_reduce --algo twoshot --dtype bf16 --out-dtype fp8 --scale 0.25 -b 4K -e 1M -f 16 --iters 3 --warmup 1
done
if [ "$NG" -ge 2 ]; then
  all=$(python3 -c "print(','.join(str(i) for i in range($NG)))")
  for mode in "" "--procs"; do
    tag=${mode:+p}
    for algo in ll ll2 oneshot twoshot nvls; do
      run g${NG}${tag}_ar_${algo} --devs $all $mode --op all_reduce --algo $algo -b 1K -e 64M -f 16 --iters 5 --warmup 2
    done
    run g${NG}${tag}_ar_auto --devs $all $mode --op all_reduce -b 1K -e 256M -f 4 --iters 5 --warmup 2
    for op in all_gather reduce_scatter; do
      for algo in ll twoshot nvls; do
        run g${NG}${tag}_${op}_${algo} --devs $all $mode --op $op --algo $algo -b 4K -e 64M -f 16 --iters 5 --warmup 2
      done
    done
    run g${NG}${tag}_a2a_ll --devs $all $mode --op alltoall --algo ll -b 4K -e 256K -f 4 --iters 5 --warmup 2
    run g${NG}${tag}_a2a_p2p --devs $all $mode --op alltoall --algo twoshot -b 4K -e 64M -f 16 --iters 5 --warmup 2
    run g${NG}${tag}_ar_nvls_f32 --devs $all $mode --op all_reduce --algo nvls --dtype f32 -b 4K -e 16M -f 16 --iters 3 --warmup 1
    run g${NG}${tag}_ar_nvls_scale --devs $all $mode --op all_reduce --algo nvls --dtype bf16 --out-dtype f32 --scale 0.5 -b 4K -e 16M -f 16 --iters 3 --warmup 1
    run g${NG}${tag}_ar_tail --devs $all $mode --op all_reduce --algo nvls -b 1030 -e 70000 -f 3 --iters 3 --warmup 1
  done
fi
echo "matrix: pass=$pass fail=$fail" | tee -a "$OUT/matrix.summary"
