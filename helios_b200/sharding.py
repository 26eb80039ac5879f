"""Multi-GPU partitioning of the RT hot path on one NVLink/NVSwitch node.  The reference is
single-GPU (it has no counterpart to anything in this file); one process drives one GPU.

Two forms, both data-parallel:

  by wavelength bin   Columns (x, y) are independent through interpolation, mixing, transmission, direct
                      beam and the flux sweeps.  Every rank owns a contiguous range of bins of every
                      bin-indexed input (`shard_store`), per-layer vectors stay replicated, and the only
                      exchange per RT iteration is the sum over ranks of the per-interface flux totals
                      (2 * ninterface doubles).  `attach_flux_allreduce` wires that exchange into
                      `Compute.integrate_flux` as ONE peer-memory kernel over NVLink
                      (csrc/comm.cu: helios_comm_allreduce_flux_totals); the sum runs in rank order on every
                      rank, so all ranks hold bitwise-identical totals and the replicated temperature step
                      cannot drift apart.
  by atmosphere       A grid of independent atmospheres (BASELINE.json configs[4]) is dealt out round-robin
                      (`partition_atmospheres`); no data-path collective at all.

torch.distributed is used for the plumbing only (exchanging the 64-byte IPC handles once, the one-time
stellar-energy sum); under the `gloo` backend the same host logic runs on CPU (tests/test_sharding_gloo.py).
"""
import numpy as np

from . import backend

# bin-indexed host inputs of a Store and their layouts (leading dims, x, trailing dims)
_BIN_VECTORS = ["opac_wave", "opac_deltawave", "surf_albedo", "starflux"]
_BIN_BY_LEVEL = ["abs_cross_all_clouds_lay", "scat_cross_all_clouds_lay", "g_0_all_clouds_lay",
                 "abs_cross_all_clouds_int", "scat_cross_all_clouds_int", "g_0_all_clouds_int"]


def bin_range(nbin, rank, world):
    """contiguous, balanced: the first nbin % world ranks hold one bin more"""
    nbin, rank, world = int(nbin), int(rank), int(world)
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, extra = divmod(nbin, world)
    x0 = rank * base + min(rank, extra)
    return x0, x0 + base + (1 if rank < extra else 0)


def _slice_x(flat, lead, nbin, trail, x0, x1):
    a = np.asarray(flat)
    if a.size != lead * nbin * trail:
        raise ValueError("array of %d elements is not [%d][%d][%d]" % (a.size, lead, nbin, trail))
    return np.ascontiguousarray(a.reshape(lead, nbin, trail)[:, x0:x1, :]).reshape(-1)


def shard_store(q, rank, world):
    """Cut every bin-indexed HOST input of an un-uploaded Store down to this rank's bins (in place) and
    re-derive the dimensions.  Call before `create_zero_arrays` / `copy_host_to_device`."""
    nbin, ny = int(q.nbin), int(q.ny)
    x0, x1 = bin_range(nbin, rank, world)
    if x1 <= x0:
        raise ValueError("rank %d of %d would own no bins (nbin = %d)" % (rank, world, nbin))
    ntp = int(q.ntemp) * int(q.npress)
    q.opac_interwave = np.ascontiguousarray(np.asarray(q.opac_interwave)[x0:x1 + 1])
    for name in _BIN_VECTORS:
        v = getattr(q, name, None)
        if v is not None and np.size(v) == nbin:
            setattr(q, name, np.ascontiguousarray(np.asarray(v)[x0:x1]))
    if q.opac_k is not None and np.size(q.opac_k) == ntp * nbin * ny:
        q.opac_k = _slice_x(q.opac_k, ntp, nbin, ny, x0, x1)
    if q.opac_scat_cross is not None and np.size(q.opac_scat_cross) == ntp * nbin:
        q.opac_scat_cross = _slice_x(q.opac_scat_cross, ntp, nbin, 1, x0, x1)
    for name in _BIN_BY_LEVEL:
        v = getattr(q, name, None)
        if v is None or np.size(v) == 0:
            continue
        levels = np.size(v) // nbin
        setattr(q, name, _slice_x(v, levels, nbin, 1, x0, x1))
    for sp in getattr(q, "species_list", []) or []:
        if sp.opacity_pretab is not None:
            sp.opacity_pretab = _slice_x(sp.opacity_pretab, ntp, nbin, ny, x0, x1)
        for name in ("scat_cross_sect_layer", "scat_cross_sect_interface"):
            v = getattr(sp, name, None)
            if v is not None:
                setattr(sp, name, _slice_x(v, np.size(v) // nbin, nbin, 1, x0, x1))
    q.nbin_global = np.int32(nbin)
    q.bin_offset = np.int32(x0)
    q.nbin = np.int32(x1 - x0)
    q.dimensions()
    return x0, x1


def exchange_handles(my_handle, rank, world, pg=None):
    """all-gather of the ranks' 64-byte IPC handles (one uint8 row per rank) through torch.distributed"""
    import torch
    import torch.distributed as dist
    mine = torch.from_numpy(np.frombuffer(bytes(my_handle), np.uint8).copy())
    if dist.get_backend(pg) == "nccl":
        mine = mine.cuda()
    rows = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(rows, mine, group=pg)
    return b"".join(bytes(r.cpu().numpy().tobytes()) for r in rows)


def attach_flux_allreduce(q, ctx, rank, world, pg=None, fused=True):
    """Create this rank's NVLink mailbox, connect the peers', and install the per-iteration exchange.

    fused (default): `helios_integrate_flux_double` itself performs the exchange -- the block that finishes the last
    interface pushes this rank's totals into the peers' mailboxes, waits for theirs and sums in rank order
    (helios_comm_set_fused; csrc/flux.cu) -- so a sharded flux solve is two launches (sweep, integration).
    fused=False: a separate one-block kernel after the integration (`q.flux_allreduce`, called by
    `Compute.integrate_flux`)."""
    import ctypes
    n = int(q.ninterface)
    handle = (ctypes.c_ubyte * 64)()
    backend._check(backend.lib().helios_comm_create(ctx.handle, int(rank), int(world), 2 * n, handle), "comm_create")
    handles = exchange_handles(handle, rank, world, pg)
    buf = (ctypes.c_ubyte * (64 * world)).from_buffer_copy(handles)
    backend._check(backend.lib().helios_comm_connect(ctx.handle, buf), "comm_connect")

    def flux_allreduce(quant):
        ctx.call("comm_allreduce_flux_totals", quant.dev_F_up_tot, quant.dev_F_down_tot, quant.dev_F_net, n)

    ctx.call("comm_set_fused", 1 if fused else 0)
    q.flux_allreduce = None if fused else flux_allreduce
    q.flux_allreduce_fused = bool(fused)
    return flux_allreduce


def correct_incident_energy_sharded(comp, q, rank, world, pg=None):
    """K:420-468 across ranks: the stellar energy sum runs over ALL bins, so the partial sums of the ranks'
    shards are added (one-time set-up, host side) before each rank rescales its own bins."""
    import torch
    import torch.distributed as dist
    if not (q.energy_correction == 1 and q.T_star > 10):
        return 1.0
    nb, dim = int(q.nbin), int(q.plancktable_dim)
    dl = np.asarray(q.opac_deltawave, np.float64)
    if q.real_star == 1:
        partial = float(np.sum(dl * q.dev_starflux.get()))
    else:
        partial = float(np.sum(dl * np.pi * q.dev_planckband_grid.view(dim * nb, nb).get()))
    t = torch.tensor([partial], dtype=torch.float64)
    if dist.get_backend(pg) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, group=pg)
    sigma = 5.6703669999999995e-5  # K:40
    corr = sigma * float(q.T_star) ** 4.0 / float(t.item())
    if q.real_star == 1:
        q.dev_starflux.set(q.dev_starflux.get() * corr)
    else:
        row = q.dev_planckband_grid.view(dim * nb, nb)
        row.set(row.get() * corr)
    return corr


def partition_atmospheres(params, rank, world):
    """deal a list of independent atmospheres out round-robin: similar cost per rank, no communication"""
    return [p for k, p in enumerate(params) if k % int(world) == int(rank)]
