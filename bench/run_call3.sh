#!/bin/bash
mkdir -p gpurun_out; export B200COLL_TIMEOUT_MS=5000
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 200 python bench.py --steps 5 --warmup 3 --table > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?" >> gpurun_out/bench_n1.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 --table > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?" >> gpurun_out/bench_n2.err
timeout 300 $TR --master-port 29522 bench.py --gpus 2 --steps 10 --warmup 3 --table --impl reference > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err; echo "rc=$?" >> gpurun_out/bench_n2_ref.err
timeout 120 ./build/b200coll_perf --devs 0,1 --op all_reduce -b 1K -e 1G --iters 20 --warmup 5 > gpurun_out/perf_n2_threads.txt 2>&1; echo "rc=$?" >> gpurun_out/perf_n2_threads.txt
tail -3 gpurun_out/smoke.log; tail -2 gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n2.err; tail -2 gpurun_out/bench_n2_ref.err; tail -2 gpurun_out/perf_n2_threads.txt
