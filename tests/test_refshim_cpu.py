"""CPU: the reference-loop harness (oracle/refshim) loads the reference's byte-compiled modules over its PyCUDA / astropy
stand-ins -- without a GPU the import must succeed (the first real use raises) -- and never touches the product."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_refshim_never_imports_the_product_or_its_library():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "oracle", "refshim")):
        for f in files:
            if not f.endswith(".py"):
                continue
            text = open(os.path.join(dirpath, f)).read()
            # the one sanctioned use: dump_host_store, which runs in the CALLER's process to make synthetic inputs
            body = text.split("def dump_host_store", 1)[0] + text.split("class _Quiet", 1)[-1] if f == "runner.py" else text
            assert not re.search(r"^\s*(from|import)\s+helios_b200", body, re.M), f
            assert "libhelios_b200" not in body.replace("Nothing here imports helios_b200 or loads libhelios_b200.so", ""), f


def test_reference_modules_load_over_the_stand_ins():
    from oracle.refshim import runner
    if not runner.available():
        pytest.skip("oracle/_ref/helios_py not built (needs /root/reference: make -C oracle ref refpy)")
    code = ("import sys; sys.path.insert(0, %r); from oracle.refshim import runner; c, q, h = runner.load_reference(); "
            "import pycuda.driver as d, astropy.constants as k; "
            "assert 'helios_b200' not in sys.modules; "
            "assert c.hsfunc is h and hasattr(c.Compute, 'radiation_loop') and hasattr(q.Store, 'allocate_on_device'); "
            "assert d.launches == 0 and k.c.cgs.value == 29979245800.0; print('ok')" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]
