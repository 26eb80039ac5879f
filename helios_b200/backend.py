"""ctypes binding of libhelios_b200.so and the device-array type that replaces PyCUDA's gpuarray.

The reference reaches the GPU through PyCUDA (`SourceModule` JIT of source/kernels.cu, computation.py:34-37;
`gpuarray.to_gpu` / `cuda.mem_alloc` / `.get()`, quantities.py:463-665).  This module is the whole
replacement for that layer: it loads the prebuilt sm_100a library, derives every prototype from
include/helios_b200.h (so the Python side cannot drift from the C-ABI), and exposes

    lib()            the loaded library (raises if it is missing -- there is NO fallback path)
    Context          one device context + stream
    DeviceArray      device buffer owned by the library, with .get() / .set() / .ptr like a gpuarray

Nothing in here computes anything on the CPU.
"""
from __future__ import annotations

import ctypes
import os
import re
import threading

import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_REPO_DIR = os.path.dirname(_PKG_DIR)
HEADER_PATH = os.path.join(_REPO_DIR, "include", "helios_b200.h")
LIB_PATH = os.path.join(_PKG_DIR, "csrc", "libhelios_b200.so")


class HeliosError(RuntimeError):
    """A libhelios_b200 call returned a non-zero status."""


class BackendMissing(ImportError):
    """The compiled CUDA library is not there.  The product path never falls back to the CPU."""


# ----------------------------------------------------------------------------- header parsing
_CTYPE = {
    "int": ctypes.c_int,
    "double": ctypes.c_double,
    "float": ctypes.c_float,
    "size_t": ctypes.c_size_t,
    "unsigned long long": ctypes.c_ulonglong,
}


def _param_ctype(decl: str):
    d = decl.replace("const ", "").strip()
    # drop the parameter name
    m = re.match(r"^(.*?[\*\s])\s*([A-Za-z_][A-Za-z_0-9]*)$", d)
    typ = (m.group(1) if m else d).strip()
    if typ.endswith("*"):
        return ctypes.c_void_p  # every pointer (device, host, handle, out-param) travels as void*
    typ = typ.replace("  ", " ")
    if typ not in _CTYPE:
        raise ValueError("unhandled C type %r in %r" % (typ, decl))
    return _CTYPE[typ]


def parse_header(path: str = HEADER_PATH):
    """-> {name: (restype_str, [param decl strings])} for every prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    text = re.sub(r"^\s*#.*$", " ", text, flags=re.M)
    protos = {}
    for m in re.finditer(r"(const char\*|int)\s+(helios_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, params = m.group(1), m.group(2), " ".join(m.group(3).split())
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        protos[name] = (ret, plist)
    return protos


_lib = None
_lib_lock = threading.Lock()
_protos = None


def lib():
    """Load libhelios_b200.so once and attach prototypes.  Raises BackendMissing if it is absent."""
    global _lib, _protos
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise BackendMissing(
                "%s not found: build it with `make -C helios_b200/csrc` (or __graft_entry__.build()). "
                "helios_b200 has no CPU fallback." % LIB_PATH
            )
        cdll = ctypes.CDLL(LIB_PATH)
        _protos = parse_header()
        for name, (ret, params) in _protos.items():
            fn = getattr(cdll, name)  # AttributeError here = header/library mismatch
            fn.restype = ctypes.c_char_p if ret == "const char*" else ctypes.c_int
            fn.argtypes = [_param_ctype(p) for p in params]
        if cdll.helios_abi_version() != 1:
            raise BackendMissing("libhelios_b200.so ABI version mismatch")
        _lib = cdll
    return _lib


def declared_symbols():
    return sorted(parse_header().keys())


def _check(rc: int, what: str):
    if rc != 0:
        msg = lib().helios_last_error()
        raise HeliosError("%s failed (status %d): %s" % (what, rc, msg.decode() if msg else "?"))


def _as_ptr(a):
    """DeviceArray | int address | None | ctypes pointer -> value for a c_void_p parameter"""
    if a is None:
        return None
    if isinstance(a, DeviceArray):
        return a.ptr
    if isinstance(a, (int, np.integer)):
        return int(a)
    return a


def _as_arg(a):
    """value for any parameter: device arrays -> pointers, numpy scalars -> Python scalars"""
    if isinstance(a, DeviceArray):
        return a.ptr
    if isinstance(a, np.integer):
        return int(a)
    if isinstance(a, np.floating):
        return float(a)
    return a


# ----------------------------------------------------------------------------- CUDA graphs
class Graph:
    """a recorded launch sequence (helios_graph_*): `with ctx.capture() as g: ...launches...`, then g.launch()"""

    def __init__(self, ctx):
        self._ctx = ctx
        self._g = ctypes.c_void_p()

    def __enter__(self):
        _check(lib().helios_graph_begin(self._ctx.handle), "helios_graph_begin")
        return self

    def __exit__(self, exc_type, exc, tb):
        rc = lib().helios_graph_end(self._ctx.handle, ctypes.byref(self._g))
        if exc_type is None:
            _check(rc, "helios_graph_end")
        return False

    def launch(self):
        _check(lib().helios_graph_launch(self._ctx.handle, self._g), "helios_graph_launch")

    def __del__(self):
        try:
            if self._g.value:
                lib().helios_graph_destroy(self._g)
                self._g = ctypes.c_void_p()
        except Exception:
            pass


# ----------------------------------------------------------------------------- context
class Context:
    """One device + one stream.  Replaces `import pycuda.autoinit` (computation.py:24)."""

    def __init__(self, device: int = 0):
        self._h = ctypes.c_void_p()
        _check(lib().helios_ctx_create(int(device), ctypes.byref(self._h)), "helios_ctx_create")
        self.device = int(device)

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().helios_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def call(self, name: str, *args):
        """Invoke helios_<name>(ctx, *args); DeviceArrays are passed as their device pointers."""
        fn = getattr(lib(), "helios_" + name)
        rc = fn(self._h, *[_as_arg(a) for a in args])
        _check(rc, "helios_" + name)

    def synchronize(self):
        _check(lib().helios_ctx_sync(self._h), "helios_ctx_sync")

    def set_stream(self, cuda_stream):
        _check(lib().helios_ctx_set_stream(self._h, ctypes.c_void_p(int(cuda_stream) if cuda_stream else 0)),
               "helios_ctx_set_stream")

    def set_fband_mode(self, mode: int):
        """0 = automatic, 1 = one thread per column, 2 = layer-parallel only"""
        _check(lib().helios_ctx_set_fband_mode(self._h, int(mode)), "helios_ctx_set_fband_mode")

    def device_info(self):
        sms = ctypes.c_int()
        l2 = ctypes.c_size_t()
        mem = ctypes.c_size_t()
        _check(lib().helios_ctx_device_info(self._h, ctypes.byref(sms), ctypes.byref(l2), ctypes.byref(mem)),
               "helios_ctx_device_info")
        return {"num_sms": sms.value, "l2_bytes": l2.value, "total_mem": mem.value}

    def launch_count(self) -> int:
        n = ctypes.c_ulonglong()
        _check(lib().helios_ctx_launch_count(self._h, ctypes.byref(n)), "helios_ctx_launch_count")
        return int(n.value)

    def bytes_allocated(self) -> int:
        n = ctypes.c_size_t()
        _check(lib().helios_ctx_bytes_allocated(self._h, ctypes.byref(n)), "helios_ctx_bytes_allocated")
        return int(n.value)

    # -- buffers
    def empty(self, shape, dtype=np.float64) -> "DeviceArray":
        return DeviceArray(self, shape, dtype)

    def zeros(self, shape, dtype=np.float64) -> "DeviceArray":
        a = DeviceArray(self, shape, dtype)
        a.fill_zero()
        return a

    def to_device(self, host) -> "DeviceArray":
        """gpuarray.to_gpu: allocate and copy."""
        h = np.ascontiguousarray(host)
        a = DeviceArray(self, h.shape, h.dtype)
        a.set(h)
        return a

    def capture(self) -> "Graph":
        return Graph(self)

    # -- events
    def event(self) -> "Event":
        return Event(self)


class Event:
    """cuda.Event replacement (computation.py:838-841)."""

    def __init__(self, ctx: Context):
        self._ctx = ctx
        self._h = ctypes.c_void_p()
        _check(lib().helios_event_create(ctx.handle, ctypes.byref(self._h)), "helios_event_create")

    def record(self):
        _check(lib().helios_event_record(self._ctx.handle, self._h), "helios_event_record")
        return self

    def synchronize(self):
        _check(lib().helios_event_synchronize(self._h), "helios_event_synchronize")
        return self

    def time_till(self, end: "Event") -> float:
        ms = ctypes.c_float()
        _check(lib().helios_event_elapsed_ms(self._h, end._h, ctypes.byref(ms)), "helios_event_elapsed_ms")
        return float(ms.value)

    def __del__(self):
        try:
            if self._h.value:
                lib().helios_event_destroy(self._h)
        except Exception:
            pass


class DeviceArray:
    """Device buffer owned by libhelios_b200 -- the gpuarray / DeviceAllocation stand-in.

    Offers what the unchanged reference files use on `dev_*` handles: `.get()` (host_functions.py:877-880),
    plus `.ptr`, `.gpudata`, `.nbytes`, `.dtype`, `.shape`, `.size`, `int(a)`.
    """

    def __init__(self, ctx: Context, shape, dtype=np.float64, _view_of=None, _ptr=None):
        self.ctx = ctx
        self.shape = (int(shape),) if np.isscalar(shape) else tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.size = int(np.prod(self.shape)) if len(self.shape) else 1
        self.nbytes = self.size * self.dtype.itemsize
        self._base = _view_of
        if _view_of is None:
            p = ctypes.c_void_p()
            _check(lib().helios_buf_alloc(ctx.handle, self.nbytes, ctypes.byref(p)), "helios_buf_alloc")
            self.ptr = int(p.value)
            self._owned = True
        else:
            self.ptr = int(_ptr)
            self._owned = False

    # pycuda spellings
    @property
    def gpudata(self):
        return self.ptr

    def __int__(self):
        return self.ptr

    def __len__(self):
        return self.shape[0] if self.shape else 1

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self.ptr, False), "version": 3,
                "strides": None}

    def view(self, offset_elems: int, shape, dtype=None) -> "DeviceArray":
        """A non-owning window into this buffer."""
        dt = np.dtype(dtype) if dtype is not None else self.dtype
        v = DeviceArray(self.ctx, shape, dt, _view_of=self, _ptr=self.ptr + offset_elems * self.dtype.itemsize)
        if (v.ptr - self.ptr) + v.nbytes > self.nbytes:
            raise ValueError("view exceeds the buffer")
        return v

    def get(self) -> np.ndarray:
        out = np.empty(self.shape, self.dtype)
        if self.nbytes:
            _check(lib().helios_buf_d2h(self.ctx.handle, out.ctypes.data_as(ctypes.c_void_p), self.ptr, self.nbytes),
                   "helios_buf_d2h")
        return out

    def set(self, host) -> "DeviceArray":
        h = np.ascontiguousarray(host, dtype=self.dtype)
        if h.size != self.size:
            raise ValueError("size mismatch: device %d vs host %d elements" % (self.size, h.size))
        if self.nbytes:
            _check(lib().helios_buf_h2d(self.ctx.handle, self.ptr, h.ctypes.data_as(ctypes.c_void_p), self.nbytes),
                   "helios_buf_h2d")
        return self

    def fill_zero(self) -> "DeviceArray":
        _check(lib().helios_buf_zero(self.ctx.handle, self.ptr, self.nbytes), "helios_buf_zero")
        return self

    def copy_from(self, other: "DeviceArray") -> "DeviceArray":
        if other.nbytes != self.nbytes:
            raise ValueError("size mismatch")
        _check(lib().helios_buf_d2d(self.ctx.handle, self.ptr, other.ptr, self.nbytes), "helios_buf_d2d")
        return self

    def free(self):
        if getattr(self, "_owned", False) and self.ptr:
            try:
                if self.ctx.handle.value:
                    lib().helios_buf_free(self.ctx.handle, ctypes.c_void_p(self.ptr))
            finally:
                self.ptr = 0
                self._owned = False

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedArray:
    """Page-locked host staging buffer (numpy view) for asynchronous H2D/D2H."""

    def __init__(self, shape, dtype=np.float64):
        self.shape = (int(shape),) if np.isscalar(shape) else tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = ctypes.c_void_p()
        _check(lib().helios_host_alloc(self.nbytes, ctypes.byref(p)), "helios_host_alloc")
        self._p = p
        buf = (ctypes.c_char * max(self.nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def h2d_async(self, ctx: Context, dst: DeviceArray):
        _check(lib().helios_buf_h2d_async(ctx.handle, dst.ptr, self._p, min(self.nbytes, dst.nbytes)),
               "helios_buf_h2d_async")

    def d2h_async(self, ctx: Context, src: DeviceArray):
        _check(lib().helios_buf_d2h_async(ctx.handle, self._p, src.ptr, min(self.nbytes, src.nbytes)),
               "helios_buf_d2h_async")

    def __del__(self):
        try:
            if self._p.value:
                self.array = None
                lib().helios_host_free(self._p)
                self._p = ctypes.c_void_p()
        except Exception:
            pass
