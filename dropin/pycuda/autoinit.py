"""`import pycuda.autoinit` (computation.py:24, host_functions.py:27).  The device context is bound lazily,
on first use, by helios_b200.runtime (device = LOCAL_RANK / HELIOS_DEVICE / 0)."""
from helios_b200 import runtime


def __getattr__(name):
    if name == "context":
        return runtime.default_context()
    if name == "device":
        return runtime.default_context().device
    raise AttributeError(name)
