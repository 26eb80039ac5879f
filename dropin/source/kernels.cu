// HELIOS B200 backend: the device code lives in the prebuilt libhelios_b200.so (helios_b200/csrc/*.cu);
// nothing compiles this file.  It exists because the unchanged source/read.py toggles the precision by
// editing "./source/kernels.cu" (read.py:170-208) and, for `precision = double`, expects the commented-out
// USE_SINGLE preamble below.  `precision = single` is not supported by this backend: Store.copy_host_to_device
// rejects it with a clear error.
/***
#define USE_SINGLE
***/
