#!/bin/bash
# 1-GPU profiling call: GPU test-suite, launch list with device times, and one `ncu --set full` capture of the N=1
# fused scale/cast copy kernel (collective kernels spin on peers, and ncu serialises launches, so multi-rank kernels cannot
# be replayed under ncu — their evidence is the SASS listing + device-timed sweeps).
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 200 python bench.py --steps 20 --warmup 5 --table > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_n1.csv python bench.py --steps 2 --warmup 3 --max 64M --no-e2e > /dev/null 2> gpurun_out/launches_n1.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_copy_scale -s 6 -c 3 -f -o gpurun_out/prof_copy_scale python bench.py --steps 2 --warmup 3 --min 256M --max 1G --no-e2e > /dev/null 2> gpurun_out/prof_copy_scale.err
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/clocks_idle.csv
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; ls -la gpurun_out/prof_copy_scale.ncu-rep 2>/dev/null
