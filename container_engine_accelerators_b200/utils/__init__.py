"""Small helpers (clock/throttle sampling for benchmark hygiene)."""
