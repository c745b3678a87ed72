#!/usr/bin/env python
"""Data-parallel training on the node's collective transport: the modern counterpart of the reference's hyper-parameter sweep jobs
(demo/gpu-training/generate_job.sh: pre-built TF images, framework-internal all-reduce). Here the gradient all-reduce goes through
`torch.distributed` backend "b200coll" (libb200coll over NVSwitch); nothing else about the training loop changes.

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 demo/gpu-training/ddp_b200coll.py --steps 50

In a pod: request the GPUs from the device plugin with GPUConfig.Transport = "b200coll" (deploy/device-plugin/gpu-config-b200coll.yaml);
Allocate then mounts /usr/local/nvidia and exports B200COLL_LIB, which is all this script needs. Synthetic data (there is no dataset
on an air-gapped node): a residual MLP on random tokens, bf16 autocast, AdamW. On a machine without GPUs it runs on CPU tensors and
the process group's Gloo fallback, which is how the unit test exercises it.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))      # run from a checkout without installing
import container_engine_accelerators_b200.parallel.process_group  # noqa: E402,F401  (registers backend "b200coll")


class Block(nn.Module):
    def __init__(self, d: int):
        super().__init__()
        self.norm = nn.LayerNorm(d)
        self.up = nn.Linear(d, 4 * d)
        self.down = nn.Linear(4 * d, d)

    def forward(self, x):
        return x + self.down(torch.nn.functional.gelu(self.up(self.norm(x))))


class Model(nn.Module):
    def __init__(self, vocab: int, d: int, layers: int):
        super().__init__()
        self.embed = nn.Embedding(vocab, d)
        self.blocks = nn.Sequential(*[Block(d) for _ in range(layers)])
        self.head = nn.Linear(d, vocab)

    def forward(self, tokens):
        return self.head(self.blocks(self.embed(tokens)))


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--batch", type=int, default=8, help="sequences per rank")
    ap.add_argument("--seq", type=int, default=512)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--vocab", type=int, default=8192)
    ap.add_argument("--backend", default="b200coll", help="b200coll (default) or nccl, for an A/B on the same box")
    ap.add_argument("--arena-pool", action="store_true", help="allocate parameters and DDP's gradient buckets in the transport's symmetric arena (zero-copy, NVLS-capable all-reduce)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29511")
    cuda = torch.cuda.is_available()
    if cuda:
        torch.cuda.set_device(local)
    device = torch.device("cuda", local) if cuda else torch.device("cpu")
    dist.init_process_group(args.backend if cuda or args.backend == "b200coll" else "gloo", rank=rank, world_size=world)
    torch.manual_seed(1234)                                       # same initial weights everywhere; DDP would broadcast them anyway
    import contextlib
    pool = contextlib.nullcontext()
    if args.arena_pool and cuda and args.backend == "b200coll":
        pool = torch.cuda.use_mem_pool(dist.group.WORLD.mem_pool())      # same allocation sequence on every rank: this is SPMD code
    with pool:
        model = Model(args.vocab, args.dim, args.layers).to(device)
        ddp = nn.parallel.DistributedDataParallel(model, device_ids=[local] if cuda else None, gradient_as_bucket_view=True)
    opt = torch.optim.AdamW(ddp.parameters(), lr=3e-4)
    gen = torch.Generator(device=device).manual_seed(100 + rank)   # each rank draws its own shard of the synthetic stream
    losses, t0 = [], None
    for step in range(args.steps):
        if step == min(3, args.steps - 1):                         # first steps warm up allocators and the communicator
            if cuda:
                torch.cuda.synchronize()
            dist.barrier()
            t0, timed_from = time.perf_counter(), step
        tokens = torch.randint(0, args.vocab, (args.batch, args.seq + 1), device=device, generator=gen)
        with torch.autocast(device_type=device.type, dtype=torch.bfloat16):
            logits = ddp(tokens[:, :-1])
            loss = nn.functional.cross_entropy(logits.float().view(-1, args.vocab), tokens[:, 1:].reshape(-1))
        opt.zero_grad(set_to_none=True)
        loss.backward()                                            # gradient buckets are all-reduced by the process group here
        opt.step()
        losses.append(loss.item())
    if cuda:
        torch.cuda.synchronize()
    dist.barrier()
    dt = time.perf_counter() - t0
    # every rank must hold the same weights after the same number of averaged-gradient steps
    probe = model.head.weight.detach().float().sum().reshape(1).clone()
    lo, hi = probe.clone(), probe.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    pg = dist.group.WORLD
    if rank == 0:
        steps = args.steps - timed_from
        print(json.dumps({"backend": dist.get_backend(), "world": world, "device": device.type, "steps_timed": steps,
                          "tokens_per_s": round(world * args.batch * args.seq * steps / dt, 1), "first_loss": round(losses[0], 4), "last_loss": round(losses[-1], 4),
                          "replicas_in_sync": bool(torch.allclose(lo, hi, rtol=1e-5, atol=1e-5)),
                          "fast_calls": getattr(pg, "fast_calls", None), "fallback_calls": getattr(pg, "fallback_calls", None)}))
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
