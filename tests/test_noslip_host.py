"""NoSlipWall (ModNoSlip.F90:44-149) restated around the boundary (rbc3d_b200/noslip.py), driven by the CPU oracle.
PARITY UNPINNED: the reference prints niter / residual per step but ships no recorded value, so the checks are the
properties the reference relies on (periodic duplicates share one unknown, GMRES reduces the residual by eps_Ewd in at
most 60 iterations, the updated tractions cancel the wall slip velocity)."""
import numpy as np
import pytest

from rbc3d_b200 import noslip, synth
from rbc3d_b200.gmres import gmres

LB = np.array([10.5, 10.5, 8.0])


def tube_case(seed=161269):
    centers = np.array([[-0.5, -0.5, 4.0], [0.5, 0.5, 1.0]]) + np.array([5.25, 5.25, 0.0])     # minit.F90:80-96
    sus = synth.make_suspension(1, nlat0=4, dealias=3, seed=seed, L=1.0, centers=centers, visc_ratio=1.0)
    sus.Lb = LB.copy()
    W = synth.make_walls(LB, [dict(radius=4.4, ntheta=20, nz=10)])
    W.f[:] = 0.0
    return sus, W


def test_v2v_and_global_vertex_numbers():
    x, e2v = synth.tube_mesh(8.0, 4.4, 20, 10, center=(5.25, 5.25))
    v2v = noslip.wall_build_v2v(x, LB)
    # ring 10 duplicates ring 0 (ModWall.F90:93-110); no other vertex has a periodic image in the mesh
    assert np.array_equal(v2v[:200], np.zeros(200, np.int32))
    assert np.array_equal(v2v[200:], np.arange(1, 21, dtype=np.int32))
    assert np.all(v2v <= np.arange(1, 221)) and np.all(v2v[v2v[v2v > 0] - 1] == 0)             # ModDataStruct.F90:171
    idx, n = noslip.indx_vert_glb([v2v, v2v])
    assert n == 400 and idx.max() == 400
    assert np.array_equal(idx[:200], np.arange(1, 201)) and np.array_equal(idx[200:220], np.arange(1, 21))
    assert np.array_equal(idx[220:420], np.arange(201, 401)) and np.array_equal(idx[420:], np.arange(201, 221))


def test_assemble_array_round_trip_and_duplicate_rule():
    _, W = tube_case()
    s = noslip.WallNoSlipSolver(W, LB, None, None, None)
    assert s.dof == 3 * 200
    rng = np.random.default_rng(3)
    u1 = rng.normal(size=s.dof)
    u = s.from_1d(u1)
    assert u.shape == (3, 220) and np.array_equal(u[:, 200:], u[:, :20])                      # duplicates read their master
    assert np.array_equal(s.to_1d(u), u1)
    # going to 1-D the loop of AssembleArray runs in vertex order, so a duplicate's value overwrites its master's
    u[:, 200:] += 1.0
    assert np.array_equal(s.from_1d(s.to_1d(u))[:, :20], u[:, 200:])


def test_gmres_zero_guess_stops_at_rtol_or_maxit():
    rng = np.random.default_rng(0)
    A = np.eye(40) + 0.3 * rng.normal(size=(40, 40)) / np.sqrt(40)
    b = rng.normal(size=40)
    x, it, hist = gmres(lambda u: A @ u, b, rtol=1e-3, maxit=60)
    assert hist[-1] < 1e-3 * hist[0] <= hist[-2] * 1.0 + 1e-300 or it == 60
    assert np.linalg.norm(A @ x - b) < 1.01e-3 * np.linalg.norm(b)
    x, it, hist = gmres(lambda u: A @ u, b, rtol=1e-30, maxit=7)
    assert it == 7 and len(hist) == 8


@pytest.mark.parametrize("with_cells", [False, True])
def test_noslip_solve_on_the_oracle(oracle_lib, with_cells):
    sus, W = tube_case()
    orc = oracle_lib.Oracle(LB)
    orc.set_cells(sus)
    orc.set_walls(W)
    orc.prepare_sing_int_on_walls()
    vbkg = np.array([0.0, 0.0, 8.0])                                                         # minicase, mtube.F90
    from oracle import harness
    rv, mv, st = harness.noslip_backend(orc, vbkg, cells=with_cells)
    s = noslip.WallNoSlipSolver(W, LB, rv, mv, st)
    slip0 = rv()
    f, niter, hist, slip = s.solve(rtol=1e-3, maxit=60)
    assert 0 < niter <= 60 and s.nmatvec == niter + (niter - 1) // 30                          # one matvec per iteration (+ restart)
    assert hist[-1] < 1e-3 * hist[0]
    assert np.all(np.diff(hist) <= 1e-12 * hist[0])                                          # GMRES residuals never grow
    # the slip velocity drops by the same factor (measured over the independent vertices, as vec_rhs is)
    r0, r1 = np.linalg.norm(s.to_1d(slip0)), np.linalg.norm(s.to_1d(slip))
    assert abs(r0 - hist[0]) < 1e-12 * r0
    assert abs(r1 - hist[-1]) < 1e-8 * r0                                                    # true residual = recurrence residual
    assert np.array_equal(f[:, 200:], f[:, :20])                                             # duplicates carry their master's traction
    # a traction that opposes the flow: mean axial traction is negative for vBkg along +z
    assert (f[2] * 1.0).mean() < 0
