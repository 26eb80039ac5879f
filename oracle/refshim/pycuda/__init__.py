"""Test-only stand-in for PyCUDA over the CUDA driver API (see ../README.md)."""
from . import driver  # noqa: F401

VERSION = (0, 0, 0)
VERSION_TEXT = "refshim"
