"""`pycuda.driver` names HELIOS uses: Context.synchronize, Event, mem_alloc (quantities.py:613-665)."""
import numpy as np

from helios_b200 import runtime


class Context(object):
    @staticmethod
    def synchronize():
        runtime.default_context().synchronize()


class Event(object):
    def __init__(self):
        self._ev = runtime.default_context().event()

    def record(self):
        self._ev.record()
        return self

    def synchronize(self):
        self._ev.synchronize()
        return self

    def time_till(self, other):
        return self._ev.time_till(other._ev)


def mem_alloc(nbytes):
    return runtime.default_context().zeros((int(nbytes) + 7) // 8, np.float64)
