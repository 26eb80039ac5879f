// Build shim for the reference oracle ("Oracle A").  It contains no reference code: it
// only #includes the reference's own source/kernels.cu from where it lies (REF_KERNELS is set by the
// Makefile) inside an extern "C" block -- exactly what PyCUDA's SourceModule does before handing the
// text to nvcc, which is why computation.py can look kernels up by their plain names.
extern "C" {
#include REF_KERNELS
}
