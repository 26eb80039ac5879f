"""GPU, 2 ranks on one node: the wavelength-sharded flux solve with the one-shot NVLink peer-memory
all-reduce (csrc/comm.cu) against the unsharded solve.  Skipped on boxes with a single GPU
(run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = dict(nbin=37, nlayer=24, ntemp=12, npress=8, plancktable_dim=700, plancktable_step=10)


def _solve(q, comp, rank=None, world=None):
    from helios_b200 import sharding
    comp.construct_planck_table(q)
    if world:
        sharding.correct_incident_energy_sharded(comp, q, rank, world)
    else:
        comp.correct_incident_energy(q)
    q.iter_value = np.int32(0)
    out = []
    for it in range(3):  # three RT iterations: the exchange is re-used, the banks alternate
        comp.interpolate_temperatures(q)
        comp.interpolate_planck(q)
        if it == 0:
            comp.interpolate_opacities_and_scattering_cross_sections(q)
            comp.interpolate_meanmolmass(q)
            if q.clouds == 1:
                comp.calc_total_g_0_of_gas_and_clouds(q)
            comp.calculate_transmission(q)
            comp.calculate_delta_z(q)
            q.delta_z_lay = q.dev_delta_z_lay.get()
            comp.hsfunc.calculate_height_z(q)
            q.dev_z_lay.set(q.z_lay)
            comp.calculate_direct_beamflux(q)
        comp.populate_spectral_flux_iteratively(q)
        comp.integrate_flux(q)
        comp.rad_temp_iteration(q)
        out.append((q.dev_F_up_tot.get(), q.dev_F_down_tot.get(), q.dev_F_net.get(), q.dev_T_lay.get()))
        q.iter_value = np.int32(it + 1)
    return out


def _store(config, ctx):
    from helios_b200 import synthetic
    q = synthetic.make_store(config, ctx=ctx, **SMALL)
    q.mu_star = np.float64(np.cos((180 - 50.0) * np.pi / 180.0))
    n = int(q.nlayer)
    q.T_lay = np.concatenate([np.linspace(2100.0, 950.0, n), [2200.0]])
    return q


def _worker(rank, world, port, config, fused, out):
    sys.path.insert(0, ROOT)
    try:
        import torch.distributed as dist
        from helios_b200 import backend, sharding, synthetic
        from helios_b200.computation import Compute
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        ctx = backend.Context(rank)
        comp = Compute(ctx, verbose=False)
        full = _solve(synthetic.upload(_store(config, ctx)), comp)
        q = _store(config, ctx)
        sharding.shard_store(q, rank, world)
        synthetic.upload(q)
        # fused: the exchange runs inside the integration kernel's epilogue; otherwise a separate one-block kernel
        sharding.attach_flux_allreduce(q, ctx, rank, world, fused=fused)
        part = _solve(q, comp, rank, world)
        worst = 0.0
        for (fu, fd, fn, T), (gu, gd, gn, gT) in zip(full, part):
            for a, b in ((fu, gu), (fd, gd), (T, gT)):
                worst = max(worst, float(np.max(np.abs(a - b) / np.maximum(np.abs(a), 1e-300))))
            assert np.array_equal(gn, gu - gd)  # F_net recomputed from the reduced totals
        assert worst < 1e-10, worst
        # bitwise agreement across ranks
        import torch
        mine = torch.from_numpy(np.concatenate([part[-1][0], part[-1][1], part[-1][3]]).copy())
        rows = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(rows, mine)
        assert all(torch.equal(r, mine) for r in rows), "ranks disagree bitwise"
        ctx.synchronize()
        backend.lib().helios_comm_destroy(ctx.handle)
        out.put((rank, "ok", worst))
        dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        import traceback
        out.put((rank, "fail", traceback.format_exc()))


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("config", ["C1", "C2"])
def test_wavelength_sharded_flux_solve_two_gpus(config, fused):
    from helios_b200 import backend
    import ctypes
    n = ctypes.c_int(0)
    backend.lib().helios_device_count(ctypes.byref(n))
    if n.value < 2:
        pytest.skip("needs 2 GPUs (have %d)" % n.value)
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mpctx = mp.get_context("spawn")
    out = mpctx.Queue()
    procs = [mpctx.Process(target=_worker, args=(r, 2, port, config, fused, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in results:
        assert status == "ok", "rank %d:\n%s" % (rank, info)
    print("\n[multi] %s (%s exchange): 2-rank wavelength-sharded solve vs unsharded, worst relative difference %.2e" %
          (config, "fused" if fused else "separate", max(r[2] for r in results)))
