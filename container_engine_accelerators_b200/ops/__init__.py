"""Bindings for the native collective library (coll/lib/libb200coll.so)."""
