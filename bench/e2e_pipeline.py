#!/usr/bin/env python
"""A/B of the end-to-end step (pinned host -> device -> all-reduce): copy then reduce on one stream (what bench.py's e2e block times
today) versus `Comm.all_reduce_from_host` (copy of chunk i+1 overlapped with the reduction of chunk i). torchrun, one rank per GPU;
device-timed, max over ranks. Prints one JSON line per size on rank 0.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29760 bench/e2e_pipeline.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from container_engine_accelerators_b200.ops import coll  # noqa: E402
from container_engine_accelerators_b200.parallel import harness  # noqa: E402


def main() -> int:
    dist = harness.Dist()
    torch.cuda.set_device(dist.local_rank)
    comm = coll.Comm.from_env(arena_mb=4096, tag="e2e") if dist.world > 1 else coll.Comm.init_all([dist.local_rank], arena_mb=4096)[0]
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for nbytes in (1 << 24, 1 << 26, 1 << 28, 1 << 30):
        n = nbytes // 2
        host = torch.full((n,), 0.25, dtype=torch.bfloat16).pin_memory()
        dev_in, dev_out = comm.empty(n, torch.bfloat16), comm.empty(n, torch.bfloat16)
        times = {}
        for name in ("sequential", "pipelined"):
            def step():
                if name == "sequential":
                    dev_in.copy_(host, non_blocking=True)
                    comm.all_reduce(dev_in, dev_out)
                else:
                    comm.all_reduce_from_host(host, dev_out)
            for _ in range(3):
                step()
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(5):
                step()
            e1.record(stream)
            torch.cuda.synchronize()
            times[name] = dist.max_([e0.elapsed_time(e1) / 5])[0]
            ok = bool(torch.all(dev_out[:1024].float() == 0.25 * dist.world).item())
            times[name + "_ok"] = ok
        if dist.rank == 0:
            print(json.dumps({"bytes": nbytes, "n_gpus": dist.world, "sequential_ms": round(times["sequential"], 3), "pipelined_ms": round(times["pipelined"], 3),
                              "speedup": round(times["sequential"] / times["pipelined"], 3), "correct": times["sequential_ok"] and times["pipelined_ok"]}), flush=True)
        comm.release(dev_out); comm.release(dev_in)
    comm.destroy()
    return 0


if __name__ == "__main__":
    sys.exit(main())
