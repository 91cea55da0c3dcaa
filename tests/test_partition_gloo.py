"""N > 1 host logic on CPU: world_size-2 gloo processes, each evaluating its share of the targets with the oracle
as the operator, summed like TargetList_CollectArray; the partition helpers are the ones bench.py and the GPU
path use."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rbc3d_b200 import partition


def test_cell_blocks_are_a_partition():
    for ncell in (1, 2, 7, 8, 64, 4096):
        for nranks in (1, 2, 3, 4, 8):
            covered = np.zeros(ncell, int)
            for r in range(nranks):
                lo, hi = partition.cell_block(ncell, nranks, r)
                covered[lo:hi] += 1
            assert (covered == 1).all()
    m = [partition.ownership_mask(8, 10, 4, r) for r in range(4)]
    assert (np.sum(m, axis=0) == 1).all()


def test_zslab_active_is_a_partition():
    rng = np.random.default_rng(0)
    Lb = np.array([10.5, 10.5, 8.0])
    x = rng.uniform(-3, 12, size=(3, 1000))         # points outside the box wrap like Fortran modulo
    for nranks in (1, 2, 3, 8):
        tot = sum(partition.zslab_active(x, Lb, nranks, r) for r in range(nranks))
        assert (tot == 1).all()


def test_cell_ownership_by_centroid_zslab():
    """whole cells by the z-slab of their centroid (what bench.py and the multi-GPU runners hand to the library as
    `active` flags): every cell on exactly one rank, the rank of DomainDecomp's slab (ModConf.F90:421-435) that holds the
    centroid, cells across the periodic boundary wrapped, and the 4096-cell lattice split evenly."""
    from rbc3d_b200 import synth
    sus = synth.make_suspension(2)
    npc = sus.nlat * sus.nlon
    for nranks in (1, 2, 3, 8):
        masks = [partition.ownership_mask_zslab(sus.x, npc, sus.Lb, nranks, r) for r in range(nranks)]
        assert (np.sum(masks, axis=0) == 1).all()
        own = partition.cell_owner_zslab(sus.x, npc, sus.Lb, nranks)
        zc = sus.x[2].reshape(-1, npc).mean(axis=1)
        assert np.array_equal(own, np.floor(zc / sus.Lb[2] * nranks).astype(int))
        for r in range(nranks):
            assert np.array_equal(masks[r].reshape(-1, npc)[:, 0], (own == r).astype(np.int32))
            assert (masks[r].reshape(-1, npc).min(axis=1) == masks[r].reshape(-1, npc).max(axis=1)).all()   # whole cells
    # a cell whose centroid lies beyond the box wraps into it
    x = sus.x.copy()
    x[2, :npc] += sus.Lb[2]
    assert np.array_equal(partition.cell_owner_zslab(x, npc, sus.Lb, 4), partition.cell_owner_zslab(sus.x, npc, sus.Lb, 4))
    # the benchmark lattice: 16 layers of 256 cells, +-0.2 jitter never crosses a slab boundary
    rng = np.random.default_rng(1)
    idx = np.stack(np.meshgrid(*[np.arange(16)] * 3, indexing="ij"), -1).reshape(-1, 3)
    L = 57.2589
    ctr = (idx + 0.5) * (L / 16) + rng.uniform(-0.2, 0.2, size=idx.shape)
    xz = np.zeros((3, 4096 * 4))
    xz[2] = np.repeat(ctr[:, 2], 4)
    for nranks in (2, 4, 8):
        own = partition.cell_owner_zslab(xz, 4, np.array([L, L, L]), nranks)
        assert np.array_equal(np.bincount(own, minlength=nranks), np.full(nranks, 4096 // nranks))


def _worker(rank, world, port, mode, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    from tests import util
    oracle.lib().orc_set_num_threads(2)
    sus = util.small_suspension(2)
    npc = sus.nlat * sus.nlon
    if mode == "cells":
        act = partition.ownership_mask(sus.ncell, npc, world, rank)
    else:
        act = partition.zslab_active(sus.x, sus.Lb, world, rank)
    act = act * (np.arange(sus.npoint) % 7 == 0)    # a subset of the targets keeps the CPU test short
    orc = oracle.Oracle(sus.Lb).set_cells(sus)
    v = orc.apply_cells(0.0, util.C2_MATVEC, orc.cell_targets(active=act.astype(np.int32)))
    t = torch.from_numpy(v)
    dist.all_reduce(t)                               # TargetList_CollectArray
    if rank == 0:
        np.save(out, t.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["cells", "zslab"])
def test_two_rank_sum_equals_single_rank(tmp_path, oracle_lib, mode):
    from tests import util
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "v.npy")
    mp.spawn(_worker, args=(2, port, mode, out), nprocs=2, join=True)
    v2 = np.load(out)
    sus = util.small_suspension(2)
    orc = oracle_lib.Oracle(sus.Lb).set_cells(sus)
    act = (np.arange(sus.npoint) % 7 == 0).astype(np.int32)
    ref = orc.apply_cells(0.0, util.C2_MATVEC, orc.cell_targets(active=act))
    assert util.rel_l2(v2, ref) < 1e-13


def _noslip_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    from rbc3d_b200 import mtube, noslip
    oracle.lib().orc_set_num_threads(2)
    sus, W = mtube.minicase_like(nlat0=6, ntheta=24, nz=12)
    act = partition.zslab_active(W.x, sus.Lb, world, rank)      # SetActiveFlag on tlist_wall
    orc = oracle.Oracle(sus.Lb).set_cells(sus)
    orc.set_walls(W)
    orc.prepare_sing_int_on_walls(active=act)                   # rows of this rank's vertices only

    def collect(v):
        t = torch.from_numpy(np.ascontiguousarray(v))
        dist.all_reduce(t)                                      # TargetList_CollectArray(tlist_wall, 3, v)
        return t.numpy()

    from oracle import harness
    s = noslip.WallNoSlipSolver(W, sus.Lb, *harness.noslip_backend(orc, mtube.VBKG, active=act, collect=collect))
    f, niter, hist, slip = s.solve()                            # GMRES runs redundantly on every rank (PETSC_COMM_SELF)
    np.savez(out % rank, f=f, niter=niter, hist=np.array(hist), slip=slip)
    dist.destroy_process_group()


def test_two_rank_wall_noslip_solve_equals_single_rank(tmp_path, oracle_lib):
    """NoSlipWall with the wall targets split by z-slab over two ranks (ModNoSlip.F90 + SetActiveFlag + CollectArray)."""
    from rbc3d_b200 import mtube, noslip
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "r%d.npz")
    mp.spawn(_noslip_worker, args=(2, port, out), nprocs=2, join=True)
    r0, r1 = np.load(out % 0), np.load(out % 1)
    assert int(r0["niter"]) == int(r1["niter"]) and np.array_equal(r0["f"], r1["f"])      # redundant solves agree bit for bit
    sus, W = mtube.minicase_like(nlat0=6, ntheta=24, nz=12)
    orc = oracle_lib.Oracle(sus.Lb).set_cells(sus)
    orc.set_walls(W)
    orc.prepare_sing_int_on_walls()
    from oracle import harness
    f, niter, hist, slip = noslip.WallNoSlipSolver(W, sus.Lb, *harness.noslip_backend(orc, mtube.VBKG)).solve()
    assert niter == int(r0["niter"])
    assert np.allclose(r0["hist"], hist, rtol=1e-9)
    assert np.linalg.norm(r0["f"] - f) < 1e-9 * np.linalg.norm(f)
    assert np.abs(r0["slip"] - slip).max() < 1e-10


@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_source_filter_is_complete_for_the_slab_targets(nranks):
    """Cell_Has_Source (ModConf.F90:467-493) keeps every cell that has a point within rc of a target of the rank's z-slab
    (real-space sum) or whose B-spline support reaches the slab's mesh planes (PME spreading), and drops the others."""
    from rbc3d_b200.ewald import SetEwaldPrms
    from tests import util
    sus = util.small_suspension(3, nlat0=4)                      # 27 small cells in a periodic box
    rc, Nb = SetEwaldPrms(sus.Lb, nranks=nranks)
    npc = sus.nlat * sus.nlon
    P = 8
    kept_any_dropped = False
    for r in range(nranks):
        dd = partition.domain_decomp(sus.Lb, rc, P, Nb[2], nranks, r)
        keep = np.array([partition.cell_has_source(sus.x[2, c * npc:(c + 1) * npc], dd, sus.Lb, nranks)
                         for c in range(sus.ncell)])
        kept_any_dropped |= not keep.all()
        act = partition.zslab_active(sus.x, sus.Lb, nranks, r).astype(bool)
        xt = sus.x[:, act]
        for c in np.nonzero(~keep)[0]:                           # a dropped cell must be out of reach of every slab target
            xs = sus.x[:, c * npc:(c + 1) * npc]
            d = xt[:, :, None] - xs[:, None, :]
            d -= np.rint(d / sus.Lb[:, None, None]) * sus.Lb[:, None, None]
            assert np.sqrt((d ** 2).sum(0)).min() > rc
            # and its B-spline support (P planes below floor(z Nb3 / L3)) misses the slab's planes
            k = np.floor(xs[2] * Nb[2] / sus.Lb[2]).astype(int)
            planes = np.mod(k[:, None] - np.arange(P)[None, :], Nb[2])
            lo, hi = r * Nb[2] // nranks, (r + 1) * Nb[2] // nranks
            assert not ((planes >= lo) & (planes < hi)).any()
        # Is_Source point-wise is implied by the cell test for every point inside the buffer
        pts = partition.is_source(sus.x[2], dd, sus.Lb, nranks)
        assert np.all(keep[np.nonzero(pts)[0] // npc])
    assert kept_any_dropped or nranks < 4                        # with 4 slabs the filter does drop cells in this box
