"""Minimal stand-in for quimb (see ../README.md)."""
from . import tensor  # noqa: F401
