"""Process-wide default device context (the stand-in for `import pycuda.autoinit`, computation.py:24).

One process drives one GPU: the device index is LOCAL_RANK when launched under torchrun, else
HELIOS_DEVICE, else 0.
"""
import os

from . import backend

_default = None


def default_device() -> int:
    # LOCAL_RANK first: under torchrun every rank must bind its own GPU even when HELIOS_DEVICE is exported
    if "LOCAL_RANK" in os.environ and "HELIOS_DEVICE" in os.environ and os.environ["LOCAL_RANK"] != os.environ["HELIOS_DEVICE"]:
        import warnings
        warnings.warn("both LOCAL_RANK and HELIOS_DEVICE are set; using LOCAL_RANK=%s" % os.environ["LOCAL_RANK"])
    for key in ("LOCAL_RANK", "HELIOS_DEVICE"):
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
