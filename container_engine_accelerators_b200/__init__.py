"""b200-node-accelerators: Blackwell-native node agent + intra-node collective library (see README.md / DESIGN.md).

Subpackages: `ops` (libb200coll binding), `parallel` (benchmark harness, stock-NCCL reference arm), `agent` (device plugin,
health, metrics, NRI injector, kube client, conformance doubles), `scheduler` (topology scheduler + labeler), `models`
(synthetic workload shapes), `utils`.
"""
__version__ = "0.1.0"
