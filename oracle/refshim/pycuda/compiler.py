"""pycuda.compiler stand-in.  PyCUDA's SourceModule(text) wraps the text in extern "C" and runs nvcc at start-up; here
that compilation was done ahead of time by `make -C oracle ref` on the same unmodified file (oracle/_ref/helios_ref.cubin),
and get_function(name)(*args, block=, grid=) is cuLaunchKernel on the NULL stream -- PyCUDA's argument conventions:
numpy scalars by value with their own width, everything int()-able as a device pointer."""
import ctypes
import os

import numpy as np

from . import driver

CUBIN = os.environ.get("REFSHIM_CUBIN", os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(
    os.path.abspath(__file__)))), "_ref", "helios_ref.cubin"))


class SourceModule(object):
    def __init__(self, source=None, **kwargs):
        driver.init()
        if not os.path.exists(CUBIN):
            raise driver.Error("reference cubin %s missing: run `make -C oracle ref` where /root/reference exists" % CUBIN)
        image = open(CUBIN, "rb").read()
        self._image = ctypes.create_string_buffer(image, len(image))
        self._mod = ctypes.c_void_p()
        driver.check(driver.cu().cuModuleLoadData(ctypes.byref(self._mod), self._image), "cuModuleLoadData")
        self._fn = {}

    def get_function(self, name):
        if name not in self._fn:
            f = ctypes.c_void_p()
            driver.check(driver.cu().cuModuleGetFunction(ctypes.byref(f), self._mod, name.encode()),
                         "cuModuleGetFunction(%s)" % name)
            self._fn[name] = _Function(name, f)
        return self._fn[name]


_SCALARS = {np.dtype(np.int32): ctypes.c_int32, np.dtype(np.int64): ctypes.c_int64, np.dtype(np.uint32): ctypes.c_uint32,
            np.dtype(np.float64): ctypes.c_double, np.dtype(np.float32): ctypes.c_float}


class _Function(object):
    def __init__(self, name, handle):
        self.name, self._f = name, handle
        self._ev = None

    def __call__(self, *args, block, grid, **kw):
        holders = []
        for a in args:
            if isinstance(a, np.generic):
                holders.append(_SCALARS[a.dtype](a.item()))
            elif a is None:
                holders.append(ctypes.c_uint64(0))
            else:
                holders.append(ctypes.c_uint64(int(a)))  # DeviceAllocation, GPUArray, raw address
        params = (ctypes.c_void_p * len(holders))(*[ctypes.cast(ctypes.byref(h), ctypes.c_void_p) for h in holders])
        g = tuple(int(v) for v in grid) + (1,) * (3 - len(grid))
        b = tuple(int(v) for v in block) + (1,) * (3 - len(block))
        cu = driver.cu()
        if driver.PROFILE:
            if self._ev is None:
                self._ev = (driver.Event(), driver.Event())
            self._ev[0].record()
        driver.check(cu.cuLaunchKernel(self._f, g[0], g[1], g[2], b[0], b[1], b[2], 0, None, params, None),
                     "cuLaunchKernel(%s)" % self.name)
        driver.launches += 1
        if driver.PROFILE:
            self._ev[1].record()
            self._ev[1].synchronize()
            rec = driver.per_kernel_ms.setdefault(self.name, [0, 0.0])
            rec[0] += 1
            rec[1] += self._ev[0].time_till(self._ev[1])
