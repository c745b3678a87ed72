"""One-rank-per-GPU launch plumbing: nccl-tests-protocol harness and the stock-NCCL reference arm."""
