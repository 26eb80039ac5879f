"""Facade for the three PyCUDA names the UNCHANGED source/host_functions.py of HELIOS touches
(host_functions.py:26-29, 919-924, 1052-1056): `import pycuda.driver`, `import pycuda.autoinit`,
`pycuda.gpuarray.to_gpu(...)` / `.get()` and `from pycuda.compiler import SourceModule`.
Device memory is owned by libhelios_b200.so; `to_gpu` returns a helios_b200 DeviceArray (which offers
`.get()`, `.ptr`, `.gpudata`, `.nbytes`, `.dtype`, `.shape`).  Put the directory that contains this package
on PYTHONPATH ahead of site-packages only where the real PyCUDA is not wanted; see INTEGRATION.md."""
VERSION = (0, 0, "helios_b200-facade")
