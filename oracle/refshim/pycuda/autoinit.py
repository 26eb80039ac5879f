"""pycuda.autoinit stand-in: importing it creates (retains) the context, as PyCUDA's does.  Without a GPU (the build
container) the import succeeds and the first real use raises instead, so that the loader can be exercised on CPU."""
from . import driver

try:
    context = driver.init()
    device = driver._device
except (OSError, driver.Error):
    context = None
    device = None
