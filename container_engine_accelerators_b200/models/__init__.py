"""There is no model code in this project (the reference has none: SURVEY §2.6, BASELINE.json "there is no model code").
What lives here are *workload shapes* that the collective benchmarks exercise — the traffic patterns the parallelism
strategies of a training/serving stack would generate (expert-parallel dispatch/combine, tensor-parallel all-reduce,
FSDP gather/scatter) — so the transport can be measured on realistic shapes without a model."""
