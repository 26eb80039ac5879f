"""pycuda.gpuarray stand-in: to_gpu / get on driver allocations."""
import numpy as np

from . import driver


class GPUArray(object):
    def __init__(self, shape, dtype, allocator=None):
        self.shape = (int(shape),) if np.isscalar(shape) else tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.size = int(np.prod(self.shape)) if len(self.shape) else 1
        self.nbytes = self.size * self.dtype.itemsize
        self.gpudata = driver.mem_alloc(self.nbytes)

    @property
    def ptr(self):
        return int(self.gpudata)

    def __int__(self):
        return int(self.gpudata)

    __index__ = __int__

    def set(self, host):
        host = np.ascontiguousarray(host, dtype=self.dtype)
        assert host.size == self.size
        driver.memcpy_htod(self.gpudata, host)
        return self

    def get(self):
        out = np.empty(self.shape, self.dtype)
        if self.nbytes:
            driver.memcpy_dtoh(out, self.gpudata)
        return out

    def __len__(self):
        return self.shape[0] if self.shape else 1


def to_gpu(host):
    host = np.ascontiguousarray(host)
    a = GPUArray(host.shape, host.dtype)
    if host.nbytes:
        driver.memcpy_htod(a.gpudata, host)
    return a


def zeros(shape, dtype=np.float64):
    a = GPUArray(shape, dtype)
    if a.nbytes:
        driver.memset_d8(a.gpudata, 0, a.nbytes)
    return a


def empty(shape, dtype=np.float64):
    return GPUArray(shape, dtype)
