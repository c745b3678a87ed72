#!/usr/bin/env python
"""Per-size timing of the end-to-end all-reduce (pinned host in -> pinned host out) through Comm.all_reduce_host, next to the plain
"copy in, all-reduce, copy back" sequence on one stream, plus the raw copy legs (H2D alone, D2H alone, both at once) so the PCIe
ceiling of the box is on record next to the number. torchrun (or plain python for one GPU); device-timed, max over ranks.
The environment knobs of coll/src/hostpath.cu (B200COLL_HOST_ZEROCOPY_KB / _PIPELINE_KB / _CHUNK_KB) are read at library load, so
A/B them by running this script once per setting:  B200COLL_HOST_CHUNK_KB=4096 python bench/e2e_hostpath.py --tag c4m"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from container_engine_accelerators_b200.ops import coll  # noqa: E402
from container_engine_accelerators_b200.parallel import harness  # noqa: E402


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--min", type=int, default=1 << 10)
    ap.add_argument("--max", type=int, default=1 << 30)
    ap.add_argument("--factor", type=int, default=4)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--tag", default="default")
    ap.add_argument("--legs", action="store_true", help="also time the raw copy legs")
    ap.add_argument("--torch-pinned", action="store_true", help="host buffers from torch.pin_memory() instead of Comm.host_empty()")
    args = ap.parse_args()
    dist = harness.Dist()
    torch.cuda.set_device(dist.local_rank)
    comm = coll.Comm.from_env(arena_mb=2304, tag="e2e") if dist.world > 1 else coll.Comm.init_all([dist.local_rank], arena_mb=2304)[0]
    n_max = args.max // 2
    if args.torch_pinned:
        h_in, h_out = torch.empty(n_max, dtype=torch.bfloat16).pin_memory(), torch.empty(n_max, dtype=torch.bfloat16).pin_memory()
    else:
        h_in, h_out = comm.host_empty(n_max, torch.bfloat16), comm.host_empty(n_max, torch.bfloat16)
    h_in.fill_(0.25)
    dev_in, dev_out = comm.empty(n_max, torch.bfloat16), comm.empty(n_max, torch.bfloat16)
    stream = torch.cuda.current_stream()
    side = torch.cuda.Stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, iters):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(iters):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return dist.max_([e0.elapsed_time(e1) / iters * 1e3])[0]

    node, cpus = comm.numa()
    if dist.rank == 0:
        print(json.dumps({"tag": args.tag, "n_gpus": dist.world, "numa_node_rank0": node, "cpus_rank0": cpus, "affinity": len(os.sched_getaffinity(0)),
                          "env": {k: v for k, v in os.environ.items() if k.startswith("B200COLL_HOST")}}), flush=True)
    nbytes = args.min
    while nbytes <= args.max:
        n = nbytes // 2
        hi, ho, di, do = h_in[:n], h_out[:n], dev_in[:n], dev_out[:n]

        def seq():
            di.copy_(hi, non_blocking=True); comm.all_reduce(di, do); ho.copy_(do, non_blocking=True)

        def lib():
            comm.all_reduce_host(hi, ho)

        def h2d():
            di.copy_(hi, non_blocking=True)

        def d2h():
            ho.copy_(do, non_blocking=True)

        def both():
            side.wait_stream(stream)
            with torch.cuda.stream(side):
                ho.copy_(do, non_blocking=True)
            di.copy_(hi, non_blocking=True)
            stream.wait_stream(side)

        row = {"bytes": nbytes, "seq_us": round(timed(seq, args.iters), 1), "lib_us": round(timed(lib, args.iters), 1)}
        ho.zero_()
        lib(); torch.cuda.synchronize()
        row["ok"] = bool((ho[:4096].float() == 0.25 * dist.world).all().item() and (ho[-4096:].float() == 0.25 * dist.world).all().item())
        row["speedup"] = round(row["seq_us"] / row["lib_us"], 3)
        row["lib_algbw"] = round(nbytes / row["lib_us"] / 1e3, 2)
        if args.legs:
            row.update({"h2d_us": round(timed(h2d, args.iters), 1), "d2h_us": round(timed(d2h, args.iters), 1), "duplex_us": round(timed(both, args.iters), 1)})
            row["h2d_gbs"] = round(nbytes / row["h2d_us"] / 1e3, 1); row["d2h_gbs"] = round(nbytes / row["d2h_us"] / 1e3, 1)
        if dist.rank == 0:
            print(json.dumps(row), flush=True)
        nbytes *= args.factor
    if dist.rank == 0:
        print(json.dumps({"stats": {k: v for k, v in comm.stats().items() if k.startswith("host")}}), flush=True)
    comm.destroy()
    dist.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
