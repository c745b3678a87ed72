"""Quick stock-NCCL sweep (torch.distributed) used by the first probe call; superseded by bench/perf.py."""
import os, sys, json, torch, torch.distributed as dist

def main():
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ops = sys.argv[1].split(",") if len(sys.argv) > 1 else ["all_reduce"]
    rows = []
    for op in ops:
        size = 1024
        while size <= (1 << 30):
            n = size // 2
            x = torch.randn(n, device="cuda").bfloat16()
            if op == "all_reduce":
                fn = lambda: dist.all_reduce(x); factor = 2 * (world - 1) / world
            elif op == "all_gather":
                out = torch.empty(n, device="cuda", dtype=torch.bfloat16); inp = x[: n // world]
                fn = lambda: dist.all_gather_into_tensor(out, inp); factor = (world - 1) / world
            elif op == "reduce_scatter":
                out = torch.empty(n // world, device="cuda", dtype=torch.bfloat16)
                fn = lambda: dist.reduce_scatter_tensor(out, x); factor = (world - 1) / world
            else:
                out = torch.empty_like(x)
                fn = lambda: dist.all_to_all_single(out, x); factor = (world - 1) / world
            iters = 50 if size <= (1 << 24) else 10
            for _ in range(5): fn()
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters): fn()
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            us = t.item() * 1e3
            algbw = size / us / 1e3
            rows.append({"op": op, "bytes": size, "us": us, "algbw": algbw, "busbw": algbw * factor})
            if rank == 0: print(f"{op:15s} {size:>12d} {us:10.2f} us  algbw {algbw:8.2f}  busbw {algbw*factor:8.2f}", flush=True)
            size *= 4
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(rows, open(f"gpurun_out/nccl_quick_n{world}.json", "w"))
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
