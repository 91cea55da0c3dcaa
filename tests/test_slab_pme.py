"""The slab-decomposed PME transform of the reference (ModPFFTW.F90: z-slabs -> all-to-all -> y-slabs and back) as an
executable NumPy specification (oracle/slabpme.py) -- SURVEY.md 8(e) (3), the multi-GPU transpose path planned for the
next round.  Checked here against the oracle's PME (single rank) and, slab-decomposed, in process and over two gloo
ranks, against the single-rank transform."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import slabpme
from tests import util

LB = np.array([3.0, 2.5, 2.0])


def point_sources(n=40, seed=3):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-0.3, 1.3, size=(3, n)) * LB[:, None]              # some outside the box: periodic wrap
    f, g, a3 = rng.normal(size=(3, n)), rng.normal(size=(3, n)), rng.normal(size=(3, n))
    return x, f, g, a3, rng.uniform(0.5, 1.5, n)


def test_bspline_and_chunks(oracle_lib):
    for P in (4, 6, 8):
        xc = np.array([0.0, 0.3, 7.999999, -2.25, 11.5])
        im, w = slabpme.bspline_func(xc, P)
        for a, i0, ww in zip(xc, im, w):
            ri, rw = oracle_lib.Oracle.bspline(a, P)
            assert ri == i0 and np.array_equal(np.asarray(rw)[:P], ww)
    assert slabpme.slab_chunks(48, 2) == [(0, 24), (24, 48)]
    assert slabpme.slab_chunks(50, 4) == [(0, 12), (12, 25), (25, 38), (38, 50)]      # ranks 1..mod get the extra plane
    assert slabpme.slab_chunks(7, 3) == [(0, 2), (2, 5), (5, 7)]


@pytest.mark.parametrize("kind", ["sl", "dl", "both"])
def test_numpy_model_equals_the_oracle_pme(oracle_lib, kind):
    orc = oracle_lib.Oracle(LB)
    m = slabpme.PmeModel(LB, orc.Nb, orc.alpha, orc.P)
    assert util.rel_l2(m.bb, orc.pme_bb().transpose(2, 1, 0)) < 1e-14
    x, f, g, a3, B = point_sources()
    c1 = 0.7 if kind in ("sl", "both") else 0.0
    c2 = -0.4 if kind in ("dl", "both") else 0.0
    orc.pme_distrib(c1, c2, x, f=f if c1 else None, g=g if c2 else None, a3=a3 if c2 else None, Bcoef=B if c2 else None)
    orc.pme_transform()
    ff, tt = m.spread(x, c1, c2, f=f, g=g, a3=a3, Bcoef=B)
    vv = m.transform(ff, tt)
    assert util.rel_l2(vv, orc.pme_vv()) < 1e-12
    xt = point_sources(25, seed=8)[0]
    v = orc.pme_interp(orc.make_targets(xt)) * 2.0                      # raw targets: Acoef = 2
    assert util.rel_l2(m.interp(xt, vv), v) < 1e-12


@pytest.mark.parametrize("R", [2, 3, 4])
def test_slab_transform_in_process(R):
    Nb = [20, 18 + (R == 4), 12 * R // np.gcd(12, R)]                   # Ny not a multiple of R for R = 4; Nz a multiple
    m = slabpme.PmeModel(LB, Nb, 0.3, 6)
    x, f, g, a3, B = point_sources()
    ff, tt = m.spread(x, 0.7, -0.4, f=f, g=g, a3=a3, Bcoef=B)
    ref = m.transform(ff, tt)
    vv = slabpme.run_slabs_in_process(m, R, ff, tt)
    assert util.rel_l2(vv, ref) < 1e-13
    # spreading restricted to a z-slab (the `cycle` of ModPME.F90:428-429) gives exactly that slab of the full mesh
    lo, hi = slabpme.slab_chunks(Nb[2], R)[1]
    fs, ts = m.spread(x, 0.7, -0.4, f=f, g=g, a3=a3, Bcoef=B, zrange=(lo, hi))
    assert np.array_equal(fs[:, lo:hi], ff[:, lo:hi]) and not fs[:, :lo].any() and not fs[:, hi:].any()
    assert np.array_equal(ts[:, lo:hi], tt[:, lo:hi])


def test_slab_interpolation_with_the_halo_of_the_lower_neighbour():
    """Interp_Vel on a z-slab + P planes from the -z neighbour (Update_Buff_Vel) = interpolation on the full mesh for the
    targets the slab owns; a target of another slab is NOT reproduced (it needs planes the rank does not hold)."""
    Nb, P, R = [20, 18, 24], 6, 3
    m = slabpme.PmeModel(LB, Nb, 0.3, P)
    rng = np.random.default_rng(4)
    vv = rng.normal(size=(3, Nb[2], Nb[1], Nb[0]))
    x = rng.uniform(-0.5, 1.5, size=(3, 400)) * LB[:, None]
    ref = m.interp(x, vv)
    kown = np.mod(np.floor(x[2] * Nb[2] / LB[2]).astype(int), Nb[2])
    for r, (lo, hi) in enumerate(slabpme.slab_chunks(Nb[2], R)):
        planes = np.arange(lo - P, hi) % Nb[2]                           # halo wraps around the period for rank 0
        mine = (kown >= lo) & (kown < hi)
        got = m.interp_slab(x[:, mine], vv[:, planes], lo, hi)
        assert mine.sum() > 50 and util.rel_l2(got, ref[:, mine]) < 1e-14
        other = m.interp_slab(x[:, ~mine], vv[:, planes], lo, hi)
        assert util.rel_l2(other, ref[:, ~mine]) > 1e-3


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Nb = [20, 19, 12]
    m = slabpme.PmeModel(LB, Nb, 0.3, 6)
    x, f, g, a3, B = point_sources()
    n = x.shape[1]
    mine = slice(rank * n // world, (rank + 1) * n // world)            # this rank's block of sources
    ff, tt = m.spread(x[:, mine], 0.7, -0.4, f=f[:, mine], g=g[:, mine], a3=a3[:, mine], Bcoef=B[mine])
    # sum over ranks, keep the own z-slab (ncclReduceScatter on the GPUs; gloo has no reduce_scatter: all_reduce + slice)
    for a in (ff, tt):
        t = torch.from_numpy(a)
        dist.all_reduce(t)
    lo, hi = slabpme.slab_chunks(Nb[2], world)[rank]

    def exchange(blocks, shapes):                                       # all-to-all out of point-to-point messages
        got = [None] * world
        got[rank] = blocks[rank]
        reqs = [dist.isend(torch.view_as_real(torch.from_numpy(np.ascontiguousarray(blocks[s]))), dst=s)
                for s in range(world) if s != rank]
        for s in range(world):
            if s != rank:
                buf = torch.zeros(list(shapes[s]) + [2], dtype=torch.float64)
                dist.recv(buf, src=s)
                got[s] = torch.view_as_complex(buf).numpy()
        for r_ in reqs:
            r_.wait()
        return got

    sr = slabpme.SlabRank(m, world, rank, exchange)
    vslab = sr.transform(ff[:, lo:hi], tt[:, lo:hi])
    gathered = [torch.zeros_like(torch.from_numpy(vslab)) for _ in range(world)]   # equal z-slabs: Nz multiple of ranks
    dist.all_gather(gathered, torch.from_numpy(np.ascontiguousarray(vslab)))
    vv = np.concatenate([t.numpy() for t in gathered], axis=1)
    if rank == 0:
        np.save(out, vv)
    dist.destroy_process_group()


def test_slab_transform_two_gloo_ranks(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "vv.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    m = slabpme.PmeModel(LB, [20, 19, 12], 0.3, 6)
    x, f, g, a3, B = point_sources()
    ref = m.transform(*m.spread(x, 0.7, -0.4, f=f, g=g, a3=a3, Bcoef=B))
    assert util.rel_l2(np.load(out), ref) < 1e-13


@pytest.mark.parametrize("R", [2, 4])
def test_slab_decomposed_operator_equals_single_rank(oracle_lib, R):
    """The reference's whole decomposition on CPU: every rank keeps the cells that pass Cell_Has_Source,
    owns the targets of its z-slab (SetActiveFlag), evaluates the real-space sums from its kept cells alone, spreads its
    kept sources onto its own mesh planes only (no mesh reduction at all), takes part in the z-slab <-> y-slab transform,
    interpolates from its slab plus the halo of the lower neighbour, and contributes its cells' share of the linear
    term (one 3-number all-reduce).  The disjoint rows of all ranks together = the single-rank operator."""
    from rbc3d_b200 import partition, synth
    from tests.util import C1_RHS
    sus = util.small_suspension(3, nlat0=4)
    orc = oracle_lib.Oracle(sus.Lb, nranks=R)                           # Nb(3) a multiple of R
    orc.set_cells(sus)
    c1, c2 = C1_RHS, -0.5 * C1_RHS
    ref = orc.apply_cells(c1, c2, orc.cell_targets())
    m = slabpme.PmeModel(sus.Lb, orc.Nb, orc.alpha, orc.P)
    npc = sus.nlat * sus.nlon
    zs = slabpme.slab_chunks(orc.Nb[2], R)
    ranks = []
    ff = np.zeros((3, orc.Nb[2], orc.Nb[1], orc.Nb[0]))
    tt = np.zeros((9,) + ff.shape[1:])
    xvint = np.zeros(3)
    for r in range(R):
        dd = partition.domain_decomp(sus.Lb, orc.rc, orc.P, orc.Nb[2], R, r)
        keep = [c for c in range(sus.ncell) if partition.cell_has_source(sus.x[2, c * npc:(c + 1) * npc], dd, sus.Lb, R)]
        sub = synth.subset(sus, keep)
        act = partition.zslab_active(sub.x, sus.Lb, R, r)
        o = oracle_lib.Oracle(sus.Lb, nranks=R).set_cells(sub)
        v = o.add_int_on_rbcs(c1, c2, o.cell_targets(active=act), flags=o.FLAG_NO_LINEAR)
        lo, hi = zs[r]
        fs, ts = m.spread(sub.x, c1, c2, f=sub.weighted(sub.f), g=sub.weighted(sub.g), a3=sub.a3,
                          Bcoef=np.repeat(sub.Bcoef, npc), zrange=(lo, hi))
        ff[:, lo:hi], tt[:, lo:hi] = fs[:, lo:hi], ts[:, lo:hi]         # the rank's own planes; nothing is summed
        own = range(*partition.cell_block(sus.ncell, R, r))              # its share of the linear term (AddLinearInt)
        for c in own:
            sl = slice(c * npc, (c + 1) * npc)
            xvint += sus.Bcoef[c] * (sus.x[:, sl] * ((sus.g[:, sl] * sus.a3[:, sl]).sum(0) * sus.dS()[sl])[None, :]).sum(1)
        ranks.append((keep, sub, act, v, lo, hi))
    xvint = -8 * np.pi * np.prod(1.0 / sus.Lb) * xvint                   # after the all-reduce
    vv = slabpme.run_slabs_in_process(m, R, ff, tt)
    total = np.zeros_like(ref)
    owned = np.zeros(sus.npoint, int)
    for keep, sub, act, v, lo, hi in ranks:
        planes = np.arange(lo - m.P, hi) % orc.Nb[2]
        on = act.astype(bool)
        A = np.repeat(sub.Acoef, npc)
        v[:, on] += m.interp_slab(sub.x[:, on], vv[:, planes], lo, hi) / A[on]
        v[:, on] += c2 * xvint[:, None] / A[on]
        glob = (np.asarray(keep)[:, None] * npc + np.arange(npc)[None, :]).reshape(-1)
        total[:, glob[on]] += v[:, on]
        owned[glob[on]] += 1
        assert not v[:, ~on].any()
    assert (owned == 1).all()
    assert util.rel_l2(total, ref) < 1e-12
