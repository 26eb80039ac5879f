"""CPU, world_size 2 over gloo: the host logic of the multi-GPU path (helios_b200/sharding.py).

Each rank cuts its wavelength shard out of the same seeded Store, runs the flux pipeline on its own bins
(with the NumPy oracle: there is no GPU here -- the device path of the same exchange is
tests/test_gpu_multi.py), and the per-interface totals are summed across ranks.  The sum must equal the
unsharded run, the IPC-handle all-gather must deliver every rank's bytes in rank order, and the atmosphere
partition must cover a grid exactly once."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from helios_b200 import sharding, synthetic  # noqa: E402

TINY = dict(nbin=11, nlayer=9, ntemp=6, npress=5, plancktable_dim=300, plancktable_step=20)


def test_bin_ranges_tile_the_spectrum():
    for nbin in (1, 7, 385, 100000):
        for world in (1, 2, 3, 8):
            if world > nbin:
                continue
            edges = [sharding.bin_range(nbin, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == nbin
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.bin_range(10, 2, 2)


def test_partition_atmospheres_covers_the_grid_once():
    grid = synthetic.grid_parameters()
    assert len(grid) == 1024
    parts = [sharding.partition_atmospheres(grid, r, 8) for r in range(8)]
    assert [len(p) for p in parts] == [128] * 8
    seen = sorted((p["T_star"], p["g"], p["table_scale"]) for part in parts for p in part)
    assert seen == sorted((p["T_star"], p["g"], p["table_scale"]) for p in grid)


def _flux_totals(q):
    from oracle.pipeline import OracleCompute, mirror_from_host
    from helios_b200 import host
    m = mirror_from_host(q)
    oc = OracleCompute()
    m.iter_value = 0
    for step in ("construct_planck_table", "interpolate_temperatures", "interpolate_planck",
                 "interpolate_opacities_and_scattering_cross_sections", "interpolate_meanmolmass"):
        getattr(oc, step)(m)
    if m.clouds == 1:
        oc.calc_total_g_0_of_gas_and_clouds(m)
    oc.calculate_transmission(m)
    oc.calculate_delta_z(m)
    m.delta_z_lay, m.z_lay, m.p_lay = m.dev_delta_z_lay, np.zeros(int(m.nlayer)), m.dev_p_lay
    host.calculate_height_z(m)
    m.dev_z_lay = m.z_lay
    oc.calculate_direct_beamflux(m)
    oc.populate_spectral_flux_iteratively(m)
    oc.integrate_flux(m)
    return m


def _store(config):
    q = synthetic.make_store(config, ctx=object(), **TINY)
    q.mu_star = np.float64(np.cos((180 - 50.0) * np.pi / 180.0))
    n = int(q.nlayer)
    q.T_lay = np.concatenate([np.linspace(2100.0, 950.0, n), [2200.0]])
    return q


def _worker(rank, world, port, config, out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. IPC-handle plumbing: 64 bytes per rank, delivered in rank order
        handle = bytes([rank * 16 + (k % 16) for k in range(64)])
        allh = sharding.exchange_handles(handle, rank, world)
        assert len(allh) == 64 * world
        for r in range(world):
            assert allh[64 * r:64 * (r + 1)] == bytes([r * 16 + (k % 16) for k in range(64)])
        # 2. wavelength sharding: partial flux totals add up to the unsharded ones
        full = _flux_totals(_store(config))
        q = _store(config)
        x0, x1 = sharding.shard_store(q, rank, world)
        assert int(q.nbin) == x1 - x0 and int(q.nbin_global) == TINY["nbin"]
        part = _flux_totals(q)
        nint, nb_full, nb = int(q.ninterface), TINY["nbin"], int(q.nbin)
        # the shard's band fluxes are the corresponding columns of the unsharded run, bit for bit
        for name in ("dev_F_up_band", "dev_F_down_band", "dev_F_dir_band"):
            want = np.asarray(getattr(full, name)).reshape(nint, nb_full)[:, x0:x1]
            assert np.array_equal(np.asarray(getattr(part, name)).reshape(nint, nb), want), name
        tot = torch.from_numpy(np.concatenate([part.dev_F_up_tot, part.dev_F_down_tot]).copy())
        dist.all_reduce(tot)
        tot = tot.numpy()
        ref = np.concatenate([full.dev_F_up_tot, full.dev_F_down_tot])
        err = float(np.max(np.abs(tot - ref) / np.maximum(np.abs(ref), 1e-300)))
        assert err < 1e-13, err
        # every rank holds the same bits (what keeps the replicated temperature step in lock-step)
        gathered = [torch.empty(tot.size, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(tot.copy()))
        assert all(np.array_equal(g.numpy(), tot) for g in gathered)
        out.put((rank, "ok", err))
    except Exception as e:  # noqa: BLE001
        import traceback
        out.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("config", ["C1", "C2"])
def test_wavelength_sharding_world2_gloo(config):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mpctx = mp.get_context("spawn")
    out = mpctx.Queue()
    procs = [mpctx.Process(target=_worker, args=(r, 2, port, config, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in results:
        assert status == "ok", "rank %d:\n%s" % (rank, info)
