"""CPU tests of the ModRepulsion restatement in the oracle (oracle/rbc3d_oracle_walls.c, last section; SURVEY.md
8(f)-4): Closest_Neighbor_Cell / Closest_Neighbor_Wall (ModRepulsion.F90:480-613) and the displacement field of
InterCellRepulsion (:270-402).  PARITY UNPINNED; pinned here by brute force over all points / triangles and by the
geometric meaning of the results (foot of a perpendicular, separation after the push)."""
import numpy as np
import pytest

from rbc3d_b200 import synth
from tests import util

EPS = 0.1     # epsDist: larger than tube.in's 0.02 so that the 0.08 gap of the test pair triggers the push


@pytest.fixture(scope="module")
def pair(oracle_lib):
    sus = util.close_pair_suspension(gap=0.08, extra=1)
    orc = oracle_lib.Oracle(sus.Lb).set_cells(sus)
    return sus, orc


def brute_other_cell(sus, i):
    npc = sus.nlat * sus.nlon
    c = i // npc
    d = sus.x - sus.x[:, [i]]
    d -= np.rint(d / sus.Lb[:, None]) * sus.Lb[:, None]
    r = np.sqrt((d ** 2).sum(0))
    r[c * npc:(c + 1) * npc] = np.inf
    return r.min(), int(r.argmin())


def test_closest_cell_matches_brute_force_and_projects(pair):
    sus, orc = pair
    npc = sus.nlat * sus.nlon
    sid = np.arange(sus.npoint) // npc + 1
    dc, xc, dw, xw = orc.closest_neighbors(sus.x, sid, EPS)
    assert np.all(np.isinf(dw))                                          # no walls set
    idx = np.concatenate([np.arange(0, 2 * npc, 7), np.arange(2 * npc, 3 * npc, 41)])
    nproj = 0
    for i in idx:
        rb, jb = brute_other_cell(sus, i)
        if rb > orc.rc:                                                  # nothing in the 27 list cells, or farther than a list cell
            assert dc[i] >= orc.rc
            continue
        if rb > 2 * EPS:
            assert dc[i] == pytest.approx(rb, rel=0, abs=1e-14)          # mesh-point distance, no refinement (:525)
        else:
            nproj += 1
            assert dc[i] <= rb + 1e-12                                   # the projection can only come closer
            # x0 is the foot of the perpendicular: xi - x0 is parallel to the neighbour's normal there
            d = sus.x[:, i] - xc[:, i]
            d -= np.rint(d / sus.Lb) * sus.Lb
            assert abs(np.linalg.norm(d) - dc[i]) < 1e-12
            nrm = sus.a3[:, jb]                                          # normal at the closest mesh point: close to the one at x0
            assert abs(abs(d @ nrm) / np.linalg.norm(d)) > 0.97
    assert nproj > 5


def test_closest_wall_matches_brute_force(oracle_lib):
    LB = np.array([10.5, 10.5, 8.0])
    centers = np.array([[5.25, 5.25, 2.0], [8.75, 5.6, 6.0]])              # the second cell close to the wall
    sus = synth.make_suspension(1, L=LB, centers=centers, seed=3)
    W = synth.make_walls(LB, [dict(radius=4.9, ntheta=36, nz=12)])
    orc = oracle_lib.Oracle(LB).set_cells(sus)
    orc.set_walls(W)
    npc = sus.nlat * sus.nlon
    sid = np.arange(sus.npoint) // npc + 1
    dc, xc, dw, xw = orc.closest_neighbors(sus.x, sid, EPS)
    e2v = W.e2v_global()
    tri = W.x[:, e2v].transpose(2, 1, 0)                                 # (NE, corner, comp)
    for i in range(npc, 2 * npc, 29):
        xi = sus.x[:, i]
        best = np.inf
        for e in range(W.NE):
            xt = tri[e, 0] + (xi - tri[e, 0]) - np.rint((xi - tri[e, 0]) / LB) * LB
            d, _, _, x0 = oracle_lib.Oracle.min_dist_to_tri(xt, tri[e])
            best = min(best, d)
        if best < orc.rc / 2:                                            # the closest triangle's centroid is in the 27 list cells
            assert dw[i] == pytest.approx(best, rel=0, abs=1e-13)
            assert abs(np.hypot(xw[0, i] - 5.25, xw[1, i] - 5.25) - 4.9) < 0.03      # a point of the (faceted) tube
    assert np.isfinite(dw[npc:]).sum() > 100
    # a wall point asks for its closest OTHER surface: its own wall is skipped (:582)
    wid = np.full(W.NV, sus.ncell + 1, dtype=np.int32)
    _, _, dww, _ = orc.closest_neighbors(W.x, wid, EPS)
    assert np.all(np.isinf(dww))


def test_inter_cell_repulsion_pushes_the_pair_apart(pair):
    sus, orc = pair
    dx, cnt, dmin = orc.inter_cell_repulsion(EPS)
    npc = sus.nlat * sus.nlon
    assert cnt > 0 and cnt == int((np.abs(dx).sum(0) > 0).sum())
    assert 0.05 < dmin <= 0.08 + 1e-9                                    # projected separation <= mesh-point gap
    moved = np.nonzero(np.abs(dx).sum(0) > 0)[0]
    assert {0, 1} <= set(np.unique(moved // npc))                        # (the random third cell sits 0.07 from cell 2)
    # every moved point goes away from the other cell by half its deficit: |dx| = (eps - rr) / 2
    sid = np.arange(sus.npoint) // npc + 1
    dc, xc, _, _ = orc.closest_neighbors(sus.x, sid, EPS)
    assert np.allclose(np.linalg.norm(dx[:, moved], axis=0), 0.5 * (EPS - dc[moved]), atol=1e-13)
    d = sus.x[:, moved] - xc[:, moved]
    assert np.all((d * dx[:, moved]).sum(0) > 0)
    # inactive points do not move (per-rank ownership, :298)
    act = np.zeros(sus.npoint, np.int32)
    act[:npc] = 1
    dx1, cnt1, _ = orc.inter_cell_repulsion(EPS, active=act)
    assert np.all(dx1[:, npc:] == 0) and np.array_equal(dx1[:, :npc], dx[:, :npc]) and 0 < cnt1 < cnt
    # after the push the pair is farther apart
    sus2 = util.close_pair_suspension(gap=0.08, extra=1)
    sus2.x = sus.x + dx
    assert util.min_gap(sus2) > util.min_gap(sus) + 0.5 * 0.015 and util.min_gap(sus2, 1, 2) > util.min_gap(sus, 1, 2)
