"""Synthetic workload shapes for the collective benchmarks (BASELINE config 4: "alltoall_perf 8xB200, expert-dispatch shape").

expert_dispatch_plan(): tokens routed top-k to experts that are sharded across ranks (expert parallelism). Returns, per
source rank, how many token rows go to each destination rank plus the row offsets both sides need — exactly the arguments of
Comm.all_to_all_v(). Routing is drawn from a Zipf-like popularity so some experts are hot (the load-imbalance case SURVEY
§7.3-7 warns about)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class DispatchPlan:
    nranks: int
    hidden: int
    rows: np.ndarray            # rows[s, d]: token rows rank s sends to rank d
    send_off: np.ndarray        # send_off[s, d]: first row of that block in s's send buffer
    recv_off: np.ndarray        # recv_off[d, s]: first row of s's block in d's recv buffer
    max_rows: int               # buffer capacity (rows) that fits every rank's send and recv side

    def bytes_sent(self, rank: int, itemsize: int = 2) -> int:
        return int(self.rows[rank].sum()) * self.hidden * itemsize


def expert_dispatch_plan(nranks: int, tokens_per_rank: int, hidden: int, experts_per_rank: int = 4, top_k: int = 2, skew: float = 1.0, seed: int = 0) -> DispatchPlan:
    rng = np.random.default_rng(seed)
    n_experts = nranks * experts_per_rank
    pop = 1.0 / np.arange(1, n_experts + 1) ** skew if skew > 0 else np.ones(n_experts)
    pop = rng.permutation(pop / pop.sum())
    rows = np.zeros((nranks, nranks), dtype=np.int64)
    for s in range(nranks):
        choice = rng.choice(n_experts, size=(tokens_per_rank, top_k), p=pop)
        dest = choice // experts_per_rank
        rows[s] = np.bincount(dest.ravel(), minlength=nranks)
    send_off = np.zeros_like(rows)
    send_off[:, 1:] = np.cumsum(rows, axis=1)[:, :-1]
    recv_off = np.zeros_like(rows)                       # recv_off[d, s]
    recv_off[:, 1:] = np.cumsum(rows.T, axis=1)[:, :-1]
    max_rows = int(max(rows.sum(axis=1).max(), rows.sum(axis=0).max()))
    return DispatchPlan(nranks, hidden, rows, send_off, recv_off, max_rows)


def uniform_plan(nranks: int, rows_per_peer: int, hidden: int) -> DispatchPlan:
    rows = np.full((nranks, nranks), rows_per_peer, dtype=np.int64)
    off = np.tile(np.arange(nranks) * rows_per_peer, (nranks, 1))
    return DispatchPlan(nranks, hidden, rows, off.copy(), off.copy(), rows_per_peer * nranks)
