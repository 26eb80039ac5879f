"""`from pycuda.compiler import SourceModule` (host_functions.py:29) is imported but never used outside the
replaced computation.py; the B200 backend ships prebuilt device code, so compiling is an error."""


class SourceModule(object):
    def __init__(self, *a, **k):
        raise RuntimeError("the B200 backend does not JIT-compile kernels.cu: device code is libhelios_b200.so")
