"""Topology-aware gang scheduler and node labeler."""
