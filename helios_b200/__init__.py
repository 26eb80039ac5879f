"""helios_b200 -- B200-native (sm_100a) backend for the radiative-transfer hot path of HELIOS.

Only what that path needs lives here: `csrc/` (CUDA kernels + C-ABI), `backend` (ctypes binding and the
gpuarray replacement), and the host-side mirrors of the reference's `quantities.Store` and
`computation.Compute`.
"""
__version__ = "0.1.0"
