"""CPU: the C-ABI library loads and exports every symbol include/helios_b200.h declares."""
import ctypes
import os
import subprocess

import pytest

from helios_b200 import backend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_declares_the_hot_path():
    protos = backend.parse_header()
    for name in ("helios_ctx_create", "helios_buf_alloc", "helios_opac_interpol", "helios_add_to_mixed_opac",
                 "helios_calc_trans_iso", "helios_calc_trans_noniso", "helios_fdir_iso", "helios_fband_iso",
                 "helios_fband_noniso", "helios_fband_matrix_iso", "helios_fband_matrix_noniso",
                 "helios_integrate_flux_double", "helios_rad_temp_iter", "helios_conv_temp_iter",
                 "helios_plancktable", "helios_comm_allreduce_sum"):
        assert name in protos, name
    assert len(protos) >= 60


def test_library_exports_every_declared_symbol():
    assert os.path.exists(backend.LIB_PATH), "build with `make -C helios_b200/csrc`"
    out = subprocess.run(["nm", "-D", "--defined-only", backend.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [s for s in backend.declared_symbols() if s not in exported]
    assert not missing, missing


def test_library_loads_and_reports_version():
    lib = backend.lib()
    assert lib.helios_abi_version() == 1
    # no GPU here: a context cannot be created, and the failure is loud and descriptive
    n = ctypes.c_int(-1)
    rc = lib.helios_device_count(ctypes.byref(n))
    if rc != 0:
        assert b"CUDA error" in lib.helios_last_error()


def test_argument_types_follow_the_header():
    protos = backend.parse_header()
    lib = backend.lib()
    ret, params = protos["helios_fband_iso"]
    assert ret == "int" and len(params) == 29
    assert len(lib.helios_fband_iso.argtypes) == 29
    assert lib.helios_fband_iso.argtypes[13] is ctypes.c_double  # g_0
    assert lib.helios_fband_iso.argtypes[-1] is ctypes.c_int  # npass


def test_no_product_module_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "helios_b200")):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_product_fails_loudly_without_the_library_or_a_gpu(tmp_path, monkeypatch):
    """no CPU fallback anywhere on the product path: a missing libhelios_b200.so raises BackendMissing on first use, and
    with the library present but no GPU the first entry point that needs the device (context creation) raises -- it does
    not silently compute on the host"""
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from helios_b200 import backend\n"
        "backend.LIB_PATH = %r\n"
        "try:\n"
        "    backend.lib()\n"
        "except backend.BackendMissing as e:\n"
        "    print('MISSING-OK', 'no CPU fallback' in str(e))\n"
        "else:\n"
        "    print('LOADED-WITHOUT-LIBRARY')\n"
    ) % (ROOT, str(tmp_path / "absent" / "libhelios_b200.so"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120).stdout
    assert "MISSING-OK True" in out, out
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        from helios_b200 import backend
        with pytest.raises(backend.HeliosError):
            backend.Context(0)
