"""pycuda.driver stand-in: device memory, events and synchronisation through libcuda (ctypes)."""
import ctypes
import os

import numpy as np

_cu = None
_ctx = None
_device = 0
launches = 0          # kernel launches made through compiler.SourceModule
per_kernel_ms = {}    # name -> [count, total ms]; filled when PROFILE is True (cuEvent per launch)
PROFILE = False


class Error(RuntimeError):
    pass


def cu():
    global _cu
    if _cu is None:
        _cu = ctypes.CDLL("libcuda.so.1")
    return _cu


def check(rc, what):
    if rc != 0:
        name = ctypes.c_char_p()
        try:
            cu().cuGetErrorName(rc, ctypes.byref(name))
        except Exception:
            pass
        raise Error("%s failed: CUresult %d (%s)" % (what, rc, name.value.decode() if name.value else "?"))


def init(device=None):
    """retain the primary context of the device (LOCAL_RANK / REFSHIM_DEVICE / 0) and make it current"""
    global _ctx, _device
    if _ctx is not None:
        check(cu().cuCtxSetCurrent(_ctx), "cuCtxSetCurrent")
        return _ctx
    if device is None:
        device = int(os.environ.get("REFSHIM_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    check(cu().cuInit(0), "cuInit")
    dev = ctypes.c_int()
    check(cu().cuDeviceGet(ctypes.byref(dev), int(device)), "cuDeviceGet")
    c = ctypes.c_void_p()
    check(cu().cuDevicePrimaryCtxRetain(ctypes.byref(c), dev), "cuDevicePrimaryCtxRetain")
    check(cu().cuCtxSetCurrent(c), "cuCtxSetCurrent")
    _ctx, _device = c, int(device)
    return _ctx


class DeviceAllocation(object):
    """what cuda.mem_alloc returns: int()-able, freed with the object"""

    def __init__(self, nbytes):
        init()
        self.nbytes = int(nbytes)
        p = ctypes.c_uint64()
        check(cu().cuMemAlloc_v2(ctypes.byref(p), max(self.nbytes, 8)), "cuMemAlloc")
        self.ptr = int(p.value)

    def __int__(self):
        return self.ptr

    __index__ = __int__

    def free(self):
        if getattr(self, "ptr", 0):
            try:
                cu().cuMemFree_v2(ctypes.c_uint64(self.ptr))
            except Exception:
                pass
            self.ptr = 0

    def __del__(self):
        self.free()


def mem_alloc(nbytes):
    return DeviceAllocation(nbytes)


def memcpy_htod(dest, src):
    src = np.ascontiguousarray(src)
    check(cu().cuMemcpyHtoD_v2(ctypes.c_uint64(int(dest)), src.ctypes.data_as(ctypes.c_void_p), src.nbytes), "cuMemcpyHtoD")


def memcpy_dtoh(dest, src):
    check(cu().cuMemcpyDtoH_v2(dest.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(int(src)), dest.nbytes), "cuMemcpyDtoH")


def memset_d8(dest, value, nbytes):
    check(cu().cuMemsetD8_v2(ctypes.c_uint64(int(dest)), ctypes.c_ubyte(value), int(nbytes)), "cuMemsetD8")


class Context(object):
    @staticmethod
    def synchronize():
        check(cu().cuCtxSynchronize(), "cuCtxSynchronize")


class Event(object):
    def __init__(self):
        init()
        self._e = ctypes.c_void_p()
        check(cu().cuEventCreate(ctypes.byref(self._e), 0), "cuEventCreate")

    def record(self, stream=None):
        check(cu().cuEventRecord(self._e, None), "cuEventRecord")
        return self

    def synchronize(self):
        check(cu().cuEventSynchronize(self._e), "cuEventSynchronize")
        return self

    def time_till(self, end):
        ms = ctypes.c_float()
        check(cu().cuEventElapsedTime(ctypes.byref(ms), self._e, end._e), "cuEventElapsedTime")
        return float(ms.value)

    def time_since(self, start):
        return start.time_till(self)

    def __del__(self):
        try:
            if self._e:
                cu().cuEventDestroy_v2(self._e)
        except Exception:
            pass
