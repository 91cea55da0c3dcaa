"""GPU parity tests: the CUDA path behind the C ABI against the CPU oracle on the same seeded inputs.

Tolerance: relative L2 error of the velocity over all targets <= 1e-10 in FP64 (BASELINE.json north_star);
cell ids and in-range neighbour sets bit-exact."""
import numpy as np
import pytest

from tests import util
from tests.util import C1_RHS, C2_MATVEC, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def sus8():
    return util.small_suspension(2)


@pytest.fixture(scope="module")
def pair8(sus8, oracle_lib):
    from rbc3d_b200.ewald import EwaldOperator
    op = EwaldOperator(sus8.Lb)
    op.set_suspension(sus8)
    orc = oracle_lib.Oracle(sus8.Lb).set_cells(sus8)
    yield op, orc
    op.close()


@pytest.fixture(scope="module")
def pair8_walk(sus8, oracle_lib):
    """the same 8 cells with the pencil-walk spreading kernel forced: a list this short takes the source-block kernel
    by default (RBC3D_SPREAD_BLOCKS: 0 = always walk, 1 = always blocks, unset = by list length)"""
    import os
    from rbc3d_b200.ewald import EwaldOperator
    old = os.environ.get("RBC3D_SPREAD_BLOCKS")
    os.environ["RBC3D_SPREAD_BLOCKS"] = "0"
    try:
        op = EwaldOperator(sus8.Lb)
        op.set_suspension(sus8)
    finally:
        if old is None:
            del os.environ["RBC3D_SPREAD_BLOCKS"]
        else:
            os.environ["RBC3D_SPREAD_BLOCKS"] = old
    orc = oracle_lib.Oracle(sus8.Lb).set_cells(sus8)
    yield op, orc
    op.close()


def test_parameters_match(pair8):
    op, orc = pair8
    assert op.rc == orc.rc
    assert op.Nb == orc.Nb


def test_cell_list_bit_exact(pair8, sus8):
    op, orc = pair8
    Nc, cid, order, start = op.cell_list()
    assert Nc == orc.Nc
    ref = orc.cell_ids(sus8.x)
    assert np.array_equal(cid, ref)
    # sorted by cell, ascending index inside a cell, offsets = prefix sums of the histogram
    assert np.array_equal(order, np.argsort(ref, kind="stable").astype(np.int32))
    assert np.array_equal(start, np.concatenate([[0], np.cumsum(np.bincount(ref, minlength=len(start) - 1))]))


def test_neighbor_sets_bit_exact(pair8, sus8):
    op, orc = pair8
    cnt, sig = op.neighbor_signature()
    rcnt, rsig = orc.neighbor_signature(sus8.x, sus8.x)
    assert np.array_equal(cnt, rcnt)
    assert np.array_equal(sig, rsig)


@pytest.mark.parametrize("c1,c2", [(0.0, C2_MATVEC), (C1_RHS, 0.0), (C1_RHS, C1_RHS)])
def test_pair_sum(pair8, c1, c2):
    op, orc = pair8
    op.set_skip_flags(1 | 2 | 4)
    v = op.AddIntOnRbcs(c1, c2)
    op.set_skip_flags(0)
    ref = orc.add_int_on_rbcs(c1, c2, orc.cell_targets(), flags=orc.FLAG_NO_SING | orc.FLAG_NO_NEARSING | orc.FLAG_NO_LINEAR)
    assert rel_l2(v, ref) < TOL


@pytest.mark.parametrize("c1,c2", [(0.0, C2_MATVEC), (C1_RHS, 0.0)])
def test_singular(pair8, c1, c2):
    op, orc = pair8
    op.set_skip_flags(2 | 4 | 8)
    v = op.AddIntOnRbcs(c1, c2)
    op.set_skip_flags(0)
    ref = orc.add_int_on_rbcs(c1, c2, orc.cell_targets(), flags=orc.FLAG_NO_PAIRS | orc.FLAG_NO_NEARSING | orc.FLAG_NO_LINEAR)
    assert rel_l2(v, ref) < TOL


def test_linear_term(pair8):
    op, orc = pair8
    op.set_skip_flags(1 | 2 | 8)
    v = op.AddIntOnRbcs(0.0, C2_MATVEC)
    op.set_skip_flags(0)
    ref = orc.add_int_on_rbcs(0.0, C2_MATVEC, orc.cell_targets(), flags=orc.FLAG_NO_PAIRS | orc.FLAG_NO_NEARSING | orc.FLAG_NO_SING)
    assert rel_l2(v, ref) < TOL


@pytest.mark.parametrize("c1,c2", [(0.0, C2_MATVEC), (C1_RHS, 0.0)])
def test_add_int_on_rbcs_full(pair8, c1, c2):
    op, orc = pair8
    v = op.AddIntOnRbcs(c1, c2)
    ref = orc.add_int_on_rbcs(c1, c2, orc.cell_targets())
    assert rel_l2(v, ref) < TOL


@pytest.mark.parametrize("spread", ["blocks", "walk"])
@pytest.mark.parametrize("c1,c2", [(0.0, C2_MATVEC), (C1_RHS, 0.0), (C1_RHS, C1_RHS)])
def test_pme_triple(request, sus8, c1, c2, spread):
    """PME_Distrib_Source -> PME_Transform -> PME_Add_Interp_Vel: mesh velocities and target velocities, with either
    spreading kernel."""
    op, orc = request.getfixturevalue("pair8" if spread == "blocks" else "pair8_walk")
    op.PME_Distrib_Source(c1, c2, cells=True)
    op.PME_Transform()
    v = op.PME_Add_Interp_Vel()
    npc = sus8.nlat * sus8.nlon
    orc.pme_distrib(c1, c2, sus8.x, sus8.weighted(sus8.f), sus8.weighted(sus8.g), sus8.a3, np.repeat(sus8.Bcoef, npc))
    orc.pme_transform()
    ref = orc.pme_interp(orc.cell_targets())
    assert rel_l2(op.pme_grid(), orc.pme_vv()) < TOL
    assert rel_l2(v, ref) < TOL


@pytest.mark.parametrize("c1,c2", [(0.0, C2_MATVEC), (C1_RHS, 0.0)])
def test_apply_matches_oracle(pair8, c1, c2):
    """the whole operator as the GMRES callbacks use it (ModVelSolver.F90:473-489, 568-582)."""
    op, orc = pair8
    v = op.apply(c1, c2)
    ref = orc.apply_cells(c1, c2, orc.cell_targets())
    assert rel_l2(v, ref) < TOL
    # resident path returns the same numbers, and accumulation into a non-zero v works
    op.apply_resident(c1, c2)
    assert rel_l2(op.get_velocity(), ref) < TOL
    v2 = op.apply(c1, c2, v=np.ones_like(ref))
    assert rel_l2(v2 - 1.0, ref) < 1e-9


def test_raw_targets(pair8, sus8):
    """CalcVelocityField-style targets (ModPostProcess.F90:40-59): indx = -1, Acoef = 2, no self mask."""
    op, orc = pair8
    rng = np.random.default_rng(3)
    xt = rng.uniform(0, sus8.Lb[0], size=(3, 500))
    # a few points outside the box and a few hugging a membrane
    xt[:, :5] -= sus8.Lb[0]
    xt[:, 5:25] = sus8.x[:, 100:2100:100] + 0.05 * sus8.a3[:, 100:2100:100]
    # inside the |dist| < 0.01*sizePat band on both sides of a membrane (jump-condition interpolation branch)
    xt[:, 25:35] = sus8.x[:, 3000:4000:100] + 0.003 * sus8.a3[:, 3000:4000:100]
    xt[:, 35:45] = sus8.x[:, 5000:6000:100] - 0.002 * sus8.a3[:, 5000:6000:100]
    xt[:, 45:50] = sus8.x[:, 7000:7500:100] - 0.2 * sus8.a3[:, 7000:7500:100]
    op.TargetList_CreateFromRaw(xt)
    from rbc3d_b200.ewald import TL_RAW
    v = op.apply(C1_RHS, C1_RHS, tlist=TL_RAW)
    ref = orc.apply_cells(C1_RHS, C1_RHS, orc.make_targets(xt))
    assert rel_l2(v, ref) < TOL


def test_inactive_targets_untouched(sus8, oracle_lib):
    from rbc3d_b200.ewald import EwaldOperator
    act = (np.arange(sus8.npoint) % 3 == 0).astype(np.int32)
    op = EwaldOperator(sus8.Lb)
    op.set_suspension(sus8, active=act)
    v = op.apply(0.0, C2_MATVEC, v=np.full((3, sus8.npoint), 7.0))
    orc = oracle_lib.Oracle(sus8.Lb).set_cells(sus8)
    ref = orc.apply_cells(0.0, C2_MATVEC, orc.cell_targets(active=act), v=np.full((3, sus8.npoint), 7.0))
    assert np.all(v[:, act == 0] == 7.0)
    assert rel_l2(v - 7.0, ref - 7.0) < TOL
    op.close()


@pytest.mark.parametrize("gap", [0.15, 0.03])
def test_near_singular_pairs(gap, oracle_lib):
    """cells almost in contact: projection, distance check, subtract, sinh re-add and the jump interpolation."""
    from rbc3d_b200.ewald import EwaldOperator
    sus = util.close_pair_suspension(gap=gap, extra=1)
    op = EwaldOperator(sus.Lb)
    op.set_suspension(sus)
    orc = oracle_lib.Oracle(sus.Lb).set_cells(sus)
    ent = op.nearsing_entries()
    assert ent["flag"].sum() > 0, "test configuration does not trigger the near-singular path"
    for c1, c2 in [(0.0, C2_MATVEC), (C1_RHS, 0.0)]:
        op.set_skip_flags(1 | 4 | 8)
        v = op.AddIntOnRbcs(c1, c2)
        op.set_skip_flags(0)
        ref = orc.add_int_on_rbcs(c1, c2, orc.cell_targets(), flags=orc.FLAG_NO_PAIRS | orc.FLAG_NO_SING | orc.FLAG_NO_LINEAR)
        assert rel_l2(v, ref) < 1e-9, (c1, c2)
        v = op.apply(c1, c2)
        ref = orc.apply_cells(c1, c2, orc.cell_targets())
        assert rel_l2(v, ref) < TOL, (c1, c2)
    op.close()


def test_noncubic_box_small_cells(oracle_lib):
    """non-cubic box (minicase-like 10.5 x 10.5 x 8), different Nb per axis, fewer modes (nlat0 = 6)."""
    from rbc3d_b200 import synth
    from rbc3d_b200.ewald import EwaldOperator
    rng = np.random.default_rng(5)
    Lb = np.array([10.5, 10.5, 8.0])
    centers = np.array([[3.0, 3.0, 2.0], [7.0, 6.5, 5.5], [3.5, 7.5, 6.0]])
    sus = synth.make_suspension(1, nlat0=6, centers=centers, L=1.0, seed=11)
    sus.Lb = Lb
    op = EwaldOperator(Lb)
    op.set_suspension(sus)
    orc = oracle_lib.Oracle(Lb).set_cells(sus)
    assert op.Nb == orc.Nb == [48, 48, 36]
    v = op.apply(C1_RHS, C1_RHS)
    ref = orc.apply_cells(C1_RHS, C1_RHS, orc.cell_targets())
    assert rel_l2(v, ref) < TOL
    op.close()


@pytest.mark.parametrize("spread", ["0", "1"])
@pytest.mark.parametrize("Nb,P", [([44, 40, 36], 8), ([12, 20, 28], 8), ([8, 8, 8], 8), ([40, 44, 36], 6), ([20, 24, 28], 4)])
def test_pme_odd_mesh_sizes_and_spline_orders(oracle_lib, monkeypatch, Nb, P, spread):
    """PME with meshes that are not multiples of the 8-cell x runs of the spreading walk (last run wraps on the right:
    scalar-reduction flush), narrower than 16 points, as small as the B-spline support (every ring slot aliases), and
    with P != 8 (the generic block kernels).  Accuracy of the Ewald split is irrelevant here: GPU and oracle use the
    same mesh.  spread = 0: pencil walks, 1: source blocks (no difference for P != 8)."""
    from rbc3d_b200 import synth
    from rbc3d_b200.ewald import EwaldOperator
    if P != 8 and spread == "1":
        pytest.skip("one spreading kernel for P != 8")
    monkeypatch.setenv("RBC3D_SPREAD_BLOCKS", spread)
    if spread == "0":
        monkeypatch.setenv("RBC3D_INTERP_DIRECT_MAX", "0")   # and the column walk instead of one warp per target
    Lb = np.array([10.5, 9.0, 8.0])
    centers = np.array([[3.0, 3.0, 2.0], [7.0, 6.5, 5.5], [9.9, 1.0, 7.6]])   # the last cell straddles three faces
    sus = synth.make_suspension(1, nlat0=6, centers=centers, L=1.0, seed=12)
    sus.Lb = Lb
    op = EwaldOperator(Lb, P=P, Nb=Nb)
    op.set_suspension(sus)
    orc = oracle_lib.Oracle(Lb, P=P, Nb=Nb).set_cells(sus)
    assert op.Nb == orc.Nb == Nb
    npc = sus.nlat * sus.nlon
    for c1, c2 in [(C1_RHS, 0.0), (0.0, C2_MATVEC), (C1_RHS, C2_MATVEC)]:
        op.PME_Distrib_Source(c1, c2, cells=True)
        op.PME_Transform()
        v = op.PME_Add_Interp_Vel()
        orc.pme_distrib(c1, c2, sus.x, sus.weighted(sus.f), sus.weighted(sus.g), sus.a3, np.repeat(sus.Bcoef, npc))
        orc.pme_transform()
        ref = orc.pme_interp(orc.cell_targets())
        assert rel_l2(op.pme_grid(), orc.pme_vv()) < TOL
        assert rel_l2(v, ref) < TOL
    op.close()


def test_timings_and_launch_count(pair8):
    op, _ = pair8
    n0 = op.launch_count()
    op.apply_resident(0.0, C2_MATVEC)
    t = op.timings()
    assert op.launch_count() > n0
    assert t["pair"] > 0 and t["sing"] > 0 and t["spread"] > 0 and t["fft"] > 0 and t["interp"] > 0


def test_singular_direct_and_cached_paths_agree(sus8, oracle_lib):
    """RBC_SingInt through the density-independent cache (default) and through the direct kernel."""
    from rbc3d_b200.ewald import EwaldOperator
    orc = oracle_lib.Oracle(sus8.Lb).set_cells(sus8)
    ref = orc.add_int_on_rbcs(0.0, C2_MATVEC, orc.cell_targets(), flags=orc.FLAG_NO_PAIRS | orc.FLAG_NO_NEARSING | orc.FLAG_NO_LINEAR)
    out = []
    for mode in (1, 0):
        op = EwaldOperator(sus8.Lb)
        op.set_sing_cache(mode)
        op.set_suspension(sus8)
        op.set_skip_flags(2 | 4 | 8)
        out.append(op.AddIntOnRbcs(0.0, C2_MATVEC))
        # a second density on the same geometry reuses the cache
        g2 = sus8.weighted(sus8.g) * 0.5
        op.SourceList_UpdateDensity(g=g2, spG=sus8.spG * 0.5)
        v2 = op.AddIntOnRbcs(0.0, C2_MATVEC)
        assert rel_l2(v2, 0.5 * ref) < TOL
        op.close()
    assert rel_l2(out[0], ref) < TOL and rel_l2(out[1], ref) < TOL
    assert rel_l2(out[0], out[1]) < 1e-12


@pytest.mark.parametrize("nlat0,dealias", [(4, 3), (8, 3), (16, 3), (12, 2)])
def test_singular_row_kernel_on_other_meshes(oracle_lib, nlat0, dealias):
    """The row-walk singular kernel on meshes other than 36 x 72: 12 x 24 (one target group of 24 lanes), 24 x 48 (two
    groups), 48 x 96 (four groups, three point streams, ONE band buffer: 129 KB bands do not fit twice), 24 x 48 with
    dealias 2 -- double layer, single layer and both, plus the deterministic sum: two applications are bit-identical."""
    from rbc3d_b200 import synth
    from rbc3d_b200.ewald import EwaldOperator
    sus = synth.make_suspension(1, nlat0=nlat0, dealias=dealias, L=9.0, seed=5,
                                centers=np.array([[2.5, 4.5, 4.5], [6.3, 4.2, 4.8]]))
    op = EwaldOperator(sus.Lb)
    op.set_suspension(sus)
    assert op.sing_cache_info()[0]
    orc = oracle_lib.Oracle(sus.Lb).set_cells(sus)
    flags = orc.FLAG_NO_PAIRS | orc.FLAG_NO_NEARSING | orc.FLAG_NO_LINEAR
    op.set_skip_flags(2 | 4 | 8)
    for c1, c2 in ((0.0, C2_MATVEC), (C1_RHS, 0.0), (C1_RHS, C1_RHS)):
        v = op.AddIntOnRbcs(c1, c2)
        ref = orc.add_int_on_rbcs(c1, c2, orc.cell_targets(), flags=flags)
        assert rel_l2(v, ref) < TOL, (nlat0, c1, c2)
        assert np.array_equal(v, op.AddIntOnRbcs(c1, c2))          # fixed summation order, no atomics
    op.close()


@pytest.mark.parametrize("c1,c2", [(0.0, C2_MATVEC), (C1_RHS, 0.0), (C1_RHS, C1_RHS)])
def test_pair_sum_dense_self_kernel_vs_cell_list(sus8, oracle_lib, c1, c2):
    """same-surface pairs through the dense per-cell kernel (default) and through the hashed cell list."""
    from rbc3d_b200.ewald import EwaldOperator
    orc = oracle_lib.Oracle(sus8.Lb).set_cells(sus8)
    ref = orc.add_int_on_rbcs(c1, c2, orc.cell_targets(), flags=orc.FLAG_NO_SING | orc.FLAG_NO_NEARSING | orc.FLAG_NO_LINEAR)
    out = []
    for mode in (3, 1, 2, 0):   # symmetric + geometry cache (default), symmetric, dense per cell, hashed cell list
        op = EwaldOperator(sus8.Lb)
        op.set_pair_self(mode)
        op.set_suspension(sus8)
        op.set_skip_flags(1 | 2 | 4)
        out.append(op.AddIntOnRbcs(c1, c2))
        op.close()
    assert all(rel_l2(o, ref) < TOL for o in out)


def test_pair_cache_partial_and_density_update(sus8, oracle_lib, monkeypatch):
    """the per-geometry coefficient cache of the same-surface double-layer pairs: only 3 of the 8 cells cached (the
    rest takes the direct kernel), and a second density on the same geometry reuses the cache."""
    from rbc3d_b200.ewald import EwaldOperator
    orc = oracle_lib.Oracle(sus8.Lb).set_cells(sus8)
    ref = orc.add_int_on_rbcs(0.0, C2_MATVEC, orc.cell_targets(), flags=orc.FLAG_NO_SING | orc.FLAG_NO_NEARSING | orc.FLAG_NO_LINEAR)
    for max_cells in ("3", "8"):
        monkeypatch.setenv("RBC3D_PAIR_CACHE_MAX_CELLS", max_cells)
        op = EwaldOperator(sus8.Lb)
        op.set_suspension(sus8)
        op.set_skip_flags(1 | 2 | 4)
        assert rel_l2(op.AddIntOnRbcs(0.0, C2_MATVEC), ref) < TOL
        op.SourceList_UpdateDensity(g=sus8.weighted(sus8.g) * 0.25, spG=sus8.spG * 0.25)
        assert rel_l2(op.AddIntOnRbcs(0.0, C2_MATVEC), 0.25 * ref) < TOL
        op.close()


def test_cell_straddling_the_periodic_boundary(oracle_lib):
    """a cell whose points are wrapped into the box individually (extent ~ L: the non-compact branch with the
    minimum-image step) and a cell shifted by a lattice vector give the same velocities as the oracle."""
    from rbc3d_b200 import synth
    from rbc3d_b200.ewald import EwaldOperator
    L = 8.0
    centers = np.array([[0.3, 4.0, 7.9], [4.2, 3.6, 4.1]])
    sus = synth.make_suspension(1, L=L, centers=centers, seed=3)
    npc = sus.nlat * sus.nlon
    x = sus.x.copy()
    x[:, :npc] = np.mod(x[:, :npc], L)            # wrap every point of cell 0 into [0, L)
    sus.x = x
    op = EwaldOperator(sus.Lb)
    op.set_suspension(sus)
    orc = oracle_lib.Oracle(sus.Lb).set_cells(sus)
    for c1, c2 in [(0.0, C2_MATVEC), (C1_RHS, 0.0)]:
        op.set_skip_flags(1 | 2 | 4)
        v = op.AddIntOnRbcs(c1, c2)
        op.set_skip_flags(0)
        ref = orc.add_int_on_rbcs(c1, c2, orc.cell_targets(), flags=orc.FLAG_NO_SING | orc.FLAG_NO_NEARSING | orc.FLAG_NO_LINEAR)
        assert rel_l2(v, ref) < TOL
    op.close()


def test_slab_decomposed_transform_on_one_rank():
    """The slab path of the PME transform (2-D transforms per plane, pack / exchange / transpose, 1-D transform in z and
    the multiplier on z-fastest pencils, per-plane Hermitian symmetrisation; pme.cu) forced on a single rank
    (RBC3D_PME_SLAB=1) against the oracle: single layer, double layer and both, on a non-cubic mesh."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from oracle import oracle; from rbc3d_b200 import synth; from rbc3d_b200.ewald import EwaldOperator\n"
        "sus = synth.make_suspension(2, L=[9.0, 10.5, 12.0])\n"
        "op = EwaldOperator(sus.Lb); op.set_suspension(sus)\n"
        "orc = oracle.Oracle(sus.Lb).set_cells(sus)\n"
        "assert len(set(op.Nb)) == 3, op.Nb\n"
        "worst = 0.0\n"
        "for c1, c2 in ((0.0, -0.0796), (0.0796, 0.0), (0.05, 0.07)):\n"
        "    v = op.apply(c1, c2); ref = orc.apply_cells(c1, c2, orc.cell_targets())\n"
        "    op.PME_Distrib_Source(c1, c2); op.PME_Transform()\n"
        "    orc.pme_distrib(c1, c2, sus.x, sus.weighted(sus.f), sus.weighted(sus.g), sus.a3, np.repeat(sus.Bcoef, sus.nlat * sus.nlon)); orc.pme_transform()\n"
        "    worst = max(worst, np.linalg.norm(v - ref) / np.linalg.norm(ref), np.linalg.norm(op.pme_grid() - orc.pme_vv()) / np.linalg.norm(orc.pme_vv()))\n"
        "print('SLAB_WORST %%.3e' %% worst)\n" % root)
    env = dict(os.environ, RBC3D_PME_SLAB="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert "SLAB_WORST" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
    worst = float(r.stdout.split("SLAB_WORST")[1].split()[0])
    assert worst < 1e-10, worst


def test_multi_gpu_parity_when_several_gpus_are_visible():
    """2-rank NCCL run of tests/run_multi_gpu.py (skipped on a single-GPU box)."""
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(root, "tests", "run_multi_gpu.py")], capture_output=True, text=True, timeout=600)
    assert "MULTI_GPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("runner,token", [("run_multi_gpu_walls.py", "MULTI_GPU_WALLS_OK"),
                                          ("run_multi_gpu_solver.py", "MULTI_GPU_SOLVER_OK")])
def test_multi_gpu_walls_and_sharded_solver_when_several_gpus_are_visible(runner, token):
    """2-rank NCCL runs of the wall paths (operators #3 / #4, no-slip solve, point-wise z-slab wall targets) and of the
    sharded cell velocity solve (skipped on a single-GPU box; run on 2 and 8 B200s by scripts/gpu_r2*.sh)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29534", os.path.join(root, "tests", runner)],
                       capture_output=True, text=True, timeout=900)
    assert token in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


# ---- density splines built on the device (Rbc_BuildSurfaceSource on the GPU) ---------------
@pytest.fixture(scope="module")
def dev_spline_op(sus8):
    from rbc3d_b200.ewald import EwaldOperator
    op = EwaldOperator(sus8.Lb)
    op.set_mesh(sus8.ncell, sus8.nlat, sus8.nlon, sus8.th, sus8.phi, sus8.w)
    op.enable_device_splines(sus8.nlat0)
    op.SourceList_UpdateCoord(sus8.x, sus8.a3, sus8.Acoef, sus8.Bcoef, sus8.area, sus8.meshSize, sus8.spx, sus8.spa3,
                              sus8.spdetj)
    op.SourceList_UpdateDensity(f=sus8.weighted(sus8.f), g=sus8.weighted(sus8.g))   # no splines passed
    yield op
    op.close()


def test_device_splines_match_host_splines(dev_spline_op, sus8):
    """ShAnalGau + ShFilter + ShSynthEqu + Spline_Build_on_Sphere as dense operators on the GPU vs the NumPy/FFT chain"""
    for which, ref in (("g", sus8.spG), ("f", sus8.spF)):
        sp = dev_spline_op.get_density_spline(which)
        for arr in range(4):     # u, u1, u2, u12 separately (different magnitudes)
            assert rel_l2(sp[:, arr], ref[:, arr]) < 1e-12, (which, arr)


@pytest.mark.parametrize("c1,c2", [(0.0, C2_MATVEC), (C1_RHS, 0.0)])
def test_operator_with_device_splines(dev_spline_op, pair8, c1, c2):
    _, orc = pair8
    ref = orc.apply_cells(c1, c2, orc.cell_targets())
    assert rel_l2(dev_spline_op.apply(c1, c2), ref) < TOL
    dev_spline_op.apply_resident(c1, c2)
    assert rel_l2(dev_spline_op.get_velocity(), ref) < TOL


def test_device_splines_follow_a_new_density(dev_spline_op, sus8, oracle_lib):
    """GMRES protocol: only g changes between matvecs (ModVelSolver.F90:560-565)"""
    import copy
    rng = np.random.default_rng(12)
    sus2 = copy.copy(sus8)
    sus2.g = sus8.g * rng.uniform(0.5, 1.5, size=(3, 1)) + 0.1 * np.roll(sus8.g, 1, axis=0)
    from rbc3d_b200 import synth
    synth.build_splines(sus2, sus8._builder, which=("G",))
    dev_spline_op.SourceList_UpdateDensity(g=sus2.weighted(sus2.g))
    assert rel_l2(dev_spline_op.get_density_spline("g"), sus2.spG) < 1e-12
    orc = oracle_lib.Oracle(sus2.Lb).set_cells(sus2)
    ref = orc.apply_cells(0.0, C2_MATVEC, orc.cell_targets())
    assert rel_l2(dev_spline_op.apply(0.0, C2_MATVEC), ref) < TOL
    dev_spline_op.SourceList_UpdateDensity(g=sus8.weighted(sus8.g))


def test_geometry_splines_built_on_the_device(sus8, pair8):
    """rbc3d_cells_set_geometry_mesh: Rbc_BuildSurfaceSource(xFlag) on the GPU (splines of x, a3, detJ from the mesh
    fields) against the host-built splines, and the whole operator on that geometry against the oracle."""
    from rbc3d_b200.ewald import EwaldOperator
    _, orc = pair8
    op = EwaldOperator(sus8.Lb)
    op.set_mesh(sus8.ncell, sus8.nlat, sus8.nlon, sus8.th, sus8.phi, sus8.w)
    op.enable_device_splines(sus8.nlat0)
    op.SourceList_UpdateCoord_mesh(sus8.x, sus8.a3, sus8.detj, sus8.Acoef, sus8.Bcoef, sus8.area, sus8.meshSize)
    for which, ref in (("x", sus8.spx), ("a3", sus8.spa3), ("detj", sus8.spdetj)):
        sp = op.get_geometry_spline(which)
        assert sp.shape == ref.shape
        assert rel_l2(sp, ref) < 1e-12, which
    op.SourceList_UpdateDensity(f=sus8.weighted(sus8.f), g=sus8.weighted(sus8.g))
    for c1, c2 in [(0.0, C2_MATVEC), (C1_RHS, 0.0)]:
        ref = orc.apply_cells(c1, c2, orc.cell_targets())
        assert rel_l2(op.apply(c1, c2), ref) < TOL
    op.close()


def test_apply_assign_equals_zero_then_accumulate(pair8):
    op, _ = pair8
    v_acc = op.apply(0.0, C2_MATVEC)                       # caller zeroes v, operator accumulates
    v_set = op.apply_assign(0.0, C2_MATVEC, v=np.full_like(v_acc, 7.0))   # previous contents are ignored
    assert rel_l2(v_set, v_acc) < 1e-13    # FP64 mesh reductions are unordered: equal up to rounding
    v_col = op.apply_collect(0.0, C2_MATVEC, v=np.full_like(v_acc, -3.0))  # one rank: CollectArray is the identity
    assert rel_l2(v_col, v_acc) < 1e-13


def test_a_second_mesh_with_a_different_cell_count_replaces_the_first(sus8, oracle_lib):
    """rbc3d_cells_set_mesh again on a live context (another ncell, another mesh size): nothing sized by the first mesh
    survives -- an apply before the new geometry is refused, and the operator on the new cells matches the oracle."""
    from rbc3d_b200 import synth
    from rbc3d_b200.ewald import EwaldOperator
    op = EwaldOperator(sus8.Lb)
    op.set_suspension(sus8)
    op.AddIntOnRbcs(0.0, C2_MATVEC)
    small = synth.make_suspension(1, nlat0=8, L=sus8.Lb, centers=np.array([[3.0, 4.0, 5.0], [7.5, 7.0, 2.0]]), seed=5)
    op.set_mesh(small.ncell, small.nlat, small.nlon, small.th, small.phi, small.w)
    with pytest.raises(Exception):
        op.AddIntOnRbcs(0.0, C2_MATVEC)
    op.set_suspension(small)
    orc = oracle_lib.Oracle(sus8.Lb).set_cells(small)
    for c1, c2 in [(0.0, C2_MATVEC), (C1_RHS, 0.0)]:
        assert rel_l2(op.AddIntOnRbcs(c1, c2), orc.add_int_on_rbcs(c1, c2, orc.cell_targets())) < TOL
    op.close()
