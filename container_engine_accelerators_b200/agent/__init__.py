"""Node agent: device plugin, MIG, sharing, health, metrics, NRI injector and their test doubles."""
