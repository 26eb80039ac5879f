"""Drive the reference's UNMODIFIED Python (source/computation.py, quantities.py, host_functions.py -- byte-compiled from
/root/reference into oracle/_ref/helios_py/source/*.bin by `make -C oracle refpy`) over the PyCUDA stand-in of this directory.
TEST INFRASTRUCTURE: used by tests/test_gpu_refloop.py and by `bench.py --impl reference`; never by the product.
Nothing here imports helios_b200 or loads libhelios_b200.so.

Host-side inputs come as a plain dict of the attributes `read.py` would have put on the Store (made by
`dump_host_store` in a separate process, see there).  Command line:
    python -m oracle.refshim.runner run <inputs.pkl> <out.npz> [--max-iter N]      one RCE run, results to out.npz
"""
import contextlib
import importlib.abc
import importlib.machinery
import importlib.util
import io
import os
import pickle
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PY = os.path.join(os.path.dirname(HERE), "_ref", "helios_py")


def available():
    return os.path.exists(os.path.join(REF_PY, "source", "computation.bin")) and \
        os.path.exists(os.path.join(os.path.dirname(HERE), "_ref", "helios_ref.cubin"))


class _ByteCodeFinder(importlib.abc.MetaPathFinder):
    """`source` and `source.X` from oracle/_ref/helios_py/source/X.bin (byte code of the unmodified reference modules)"""

    def find_spec(self, fullname, path=None, target=None):
        if fullname != "source" and not fullname.startswith("source."):
            return None
        pkg = os.path.join(REF_PY, "source")
        leaf = "__init__" if fullname == "source" else fullname.split(".", 1)[1]
        if "." in leaf:
            return None
        f = os.path.join(pkg, leaf + ".bin")
        if not os.path.exists(f):
            return None
        loader = importlib.machinery.SourcelessFileLoader(fullname, f)
        return importlib.util.spec_from_file_location(fullname, f, loader=loader,
                                                      submodule_search_locations=[pkg] if fullname == "source" else None)


def load_reference():
    """import the reference's modules over the stand-ins; returns (computation, quantities, host_functions)"""
    for name in ("pycuda", "astropy", "source"):
        mod = sys.modules.get(name)
        if mod is not None and not (getattr(mod, "__file__", "") or "").startswith((HERE, REF_PY)):
            raise RuntimeError("a foreign `%s` is already imported (%s): run the reference in its own process" %
                               (name, getattr(mod, "__file__", "?")))
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    if not any(isinstance(f, _ByteCodeFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _ByteCodeFinder())
    import source.computation as computation
    import source.quantities as quantities
    import source.host_functions as host_functions
    return computation, quantities, host_functions


def dump_host_store(config, path, **kw):
    """(runs in the CALLER's process, which may import the product's synthetic-input generator) write the host-side
    attributes of a synthetic Store as a pickled dict"""
    from helios_b200 import synthetic
    q = synthetic.make_store(config, **kw)
    d = {}
    for k, v in vars(q).items():
        if k.startswith("dev_") or k.startswith("_") or isinstance(v, (types.FunctionType, types.MethodType)):
            continue
        if k == "species_list":
            v = [{kk: vv for kk, vv in vars(sp).items()} for sp in v]
        d[k] = v
    with open(path, "wb") as f:
        pickle.dump(d, f, protocol=4)
    return d


class _Quiet(object):
    """write / read / rt_plot arguments of the loops: the run never plots, couples or aborts in these tests"""

    def __getattr__(self, name):
        def nop(*a, **k):
            return None
        return nop


class RefRun(object):
    """the call sequence of helios.py:76-95 on a reference Store filled from `host` (dict)"""

    def __init__(self, host, quiet=True):
        self.computation, self.quantities, self.hsfunc = load_reference()
        q = self.quantities.Store()
        for k, v in host.items():
            if k == "species_list":
                v = [types.SimpleNamespace(**sp) for sp in v]
            setattr(q, k, v)
        q.dimensions()
        q.create_zero_arrays()
        q.convert_input_list_to_array()
        q.copy_host_to_device()
        q.allocate_on_device()
        self.q = q
        cwd = os.getcwd()
        os.chdir(REF_PY)  # Compute() opens ./source/kernels.cu (C:34-37); the stand-in loads the prebuilt cubin instead
        try:
            self.comp = self.computation.Compute()
        finally:
            os.chdir(cwd)
        self.quiet = quiet
        self.log = io.StringIO()

    @contextlib.contextmanager
    def _out(self):
        if self.quiet:
            with contextlib.redirect_stdout(self.log):
                yield
        else:
            yield

    def setup(self):
        with self._out():
            self.comp.construct_planck_table(self.q)
            self.comp.correct_incident_energy(self.q)

    def rce(self):
        """radiation_loop + convection_loop to the reference's own convergence criterion; returns a result dict"""
        from pycuda import driver
        q, comp = self.q, self.comp
        nul = _Quiet()
        driver.Context.synchronize()
        n0 = driver.launches
        t0 = time.perf_counter()
        status, rad_iters, conv_iters = "converged", 0, 0
        try:
            with self._out():
                comp.radiation_loop(q, nul, nul, nul)
                rad_iters = int(q.iter_value)
                if q.convection == 1:
                    comp.convection_loop(q, nul, nul, nul)
                    conv_iters = int(q.iter_value)
        except SystemExit:
            status = "iteration limit"
            rad_iters = rad_iters or int(q.iter_value)
        driver.Context.synchronize()
        dt = time.perf_counter() - t0
        return {"seconds": dt, "status": status, "radiation_iterations": rad_iters, "convection_iterations": conv_iters,
                "launches": driver.launches - n0, "T_lay": q.dev_T_lay.get(), "F_up_band": q.dev_F_up_band.get(),
                "F_net": q.dev_F_net.get(), "F_down_tot": q.dev_F_down_tot.get(), "F_up_tot": q.dev_F_up_tot.get(),
                "conv_layer": np.asarray(getattr(q, "conv_layer", np.zeros(1)), np.int32)}

    def prepare_flux_solve(self, T_lay=None):
        """everything one flux solve needs (C:856-879) for the profile T_lay"""
        q, comp, hs = self.q, self.comp, self.hsfunc
        if T_lay is not None:
            from pycuda import gpuarray
            q.T_lay = np.asarray(T_lay, np.float64)
            q.dev_T_lay = gpuarray.to_gpu(q.T_lay)
        q.iter_value = np.int32(0)
        with self._out():
            comp.interpolate_temperatures(q)
            comp.interpolate_planck(q)
            if q.opacity_mixing == "premixed":
                comp.interpolate_opacities_and_scattering_cross_sections(q)
                comp.interpolate_meanmolmass(q)
            else:
                hs.calculate_vmr_for_all_species(q)
                hs.calculate_meanmolecularmass(q)
                hs.nullify_opac_scat_arrays(q)
                comp.calculate_total_opacity_and_scat_cross_sections_from_species(q)
            if q.clouds == 1:
                comp.calc_total_g_0_of_gas_and_clouds(q)
            comp.calculate_transmission(q)
            comp.calculate_delta_z(q)
            q.delta_z_lay = q.dev_delta_z_lay.get()
            hs.calculate_height_z(q)
            from pycuda import gpuarray
            q.dev_z_lay = gpuarray.to_gpu(q.z_lay)
            comp.calculate_direct_beamflux(q)

    def flux_solve(self):
        """C:881-888: the reference's own wrapper methods, one device sync after every launch as it does"""
        self.comp.populate_spectral_flux_iteratively(self.q)
        self.comp.integrate_flux(self.q)


def main(argv):
    if len(argv) >= 3 and argv[0] == "run":
        host = pickle.load(open(argv[1], "rb"))
        if "--max-iter" in argv:
            host["max_nr_iterations"] = int(argv[argv.index("--max-iter") + 1])
        run = RefRun(host)
        run.setup()
        res = run.rce()
        np.savez(argv[2], **{k: v for k, v in res.items()})
        print("reference rce: %s after %d + %d iterations, %.2f s, %d launches" %
              (res["status"], res["radiation_iterations"], res["convection_iterations"], res["seconds"], res["launches"]))
        return 0
    print(__doc__)
    return 2


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
