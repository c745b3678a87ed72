#!/usr/bin/env python
"""alltoallv_perf — expert-dispatch shaped all-to-all (BASELINE config 4), libb200coll vs NCCL on the same box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 bench/alltoallv_perf.py

For each (tokens per rank, hidden, skew): device-timed dispatch of bf16 token rows to the ranks that own their experts, with
(a) libb200coll all_to_all_v, bf16 on the wire; (b) the same with the fp8-e4m3 quantise fused into the dispatch kernel (half the
bytes on NVLink, no separate cast kernel); (c) NCCL all_to_all_single with split sizes (torch.distributed, the way MoE stacks call it).
Bus bandwidth = (bytes a rank sends off-chip, max over ranks) / time. Rank 0 prints one JSON line per shape."""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main() -> int:
    import torch
    import torch.distributed as dist
    from container_engine_accelerators_b200.models.workloads import expert_dispatch_plan
    from container_engine_accelerators_b200.ops import coll
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    iters, warm = int(os.environ.get("ITERS", 20)), 5
    shapes = [(4096, 4096, 0.0), (4096, 7168, 1.0), (16384, 7168, 1.0), (65536, 7168, 1.2)]
    cap_rows = max(expert_dispatch_plan(world, t, h, skew=s).max_rows for t, h, s in shapes)
    cap_hidden = max(h for _, h, _ in shapes)
    comm = coll.Comm.from_env(arena_mb=int(cap_rows * cap_hidden * 2 * 2.6 / (1 << 20)) + 256, tag="a2av")
    send = comm.empty(cap_rows * cap_hidden, torch.bfloat16)
    recv = comm.empty(cap_rows * cap_hidden, torch.bfloat16)
    recv8 = comm.empty(cap_rows * cap_hidden, torch.float8_e4m3fn)
    send.copy_((torch.arange(send.numel(), device="cuda") % 251).to(torch.bfloat16) * 0.01)
    nccl_recv = torch.empty(cap_rows * cap_hidden, dtype=torch.bfloat16, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn) -> float:
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() * 1e3
    for tokens, hidden, skew in shapes:
        plan = expert_dispatch_plan(world, tokens, hidden, skew=skew, seed=1)
        rows, so, ro = plan.rows[rank].tolist(), plan.send_off[rank].tolist(), [int(plan.recv_off[d, rank]) for d in range(world)]
        n_send = int(plan.rows[rank].sum()) * hidden
        n_recv = int(plan.rows[:, rank].sum()) * hidden
        us_ours = timed(lambda: comm.all_to_all_v(send, recv, hidden, rows, so, ro))
        comm.check_async_error()
        # correctness of the bf16 path against NCCL's result
        in_split = [r * hidden for r in rows]
        out_split = [int(plan.rows[s, rank]) * hidden for s in range(world)]
        dist.all_to_all_single(nccl_recv[:n_recv], send[:n_send], out_split, in_split)
        torch.cuda.synchronize()
        ok = bool(torch.equal(recv[:n_recv], nccl_recv[:n_recv]))
        us_fp8 = timed(lambda: comm.all_to_all_v(send, recv8, hidden, rows, so, ro, scale=1.0))
        us_nccl = timed(lambda: dist.all_to_all_single(nccl_recv[:n_recv], send[:n_send], out_split, in_split))
        offchip = max(int(plan.rows[s].sum() - plan.rows[s, s]) for s in range(world)) * hidden * 2
        oks = torch.tensor([1.0 if ok else 0.0], device="cuda"); dist.all_reduce(oks, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(json.dumps({"bench": "alltoallv_perf", "n_gpus": world, "tokens_per_rank": tokens, "hidden": hidden, "skew": skew, "top_k": 2,
                              "max_offchip_bytes_per_rank": offchip, "imbalance_max_over_mean": round(float(plan.rows.sum(axis=0).max() / plan.rows.sum(axis=0).mean()), 3),
                              "ours_us": round(us_ours, 2), "ours_busbw": round(offchip / us_ours / 1e3, 2), "ours_fused_fp8_us": round(us_fp8, 2),
                              "nccl_us": round(us_nccl, 2), "nccl_busbw": round(offchip / us_nccl / 1e3, 2), "speedup": round(us_nccl / us_ours, 2),
                              "fp8_speedup_vs_nccl_bf16": round(us_nccl / us_fp8, 2), "matches_nccl": bool(oks.item() == 1.0)}), flush=True)
    comm.destroy()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
