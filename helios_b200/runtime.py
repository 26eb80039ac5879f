"""Process-wide default device context (the stand-in for `import pycuda.autoinit`, computation.py:24).

One process drives one GPU: the device index is LOCAL_RANK when launched under torchrun, else
HELIOS_DEVICE, else 0.
"""
import os

from . import backend

_default = None


def default_device() -> int:
    for key in ("HELIOS_DEVICE", "LOCAL_RANK"):
        if key in os.environ:
            return int(os.environ[key])
    return 0


def default_context() -> backend.Context:
    global _default
    if _default is None:
        _default = backend.Context(default_device())
    return _default


def set_default_context(ctx: backend.Context):
    global _default
    _default = ctx
    return ctx
