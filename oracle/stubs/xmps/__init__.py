"""Minimal stand-in for xmps @ 0fe1c7ec (see ../README.md)."""
