"""The reference's own input files through the harness readers (rbc3d_b200/cases.py) and the oracle: examples/minicase
(BASELINE.json configs[0]).  CPU only; skipped where /root/reference is not mounted (the GPU box)."""
import os

import numpy as np
import pytest

REF = "/root/reference/examples/minicase/Input"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")


def test_tube_in_and_wall_mesh_readers():
    from rbc3d_b200 import cases
    cfg = cases.read_tube_in(os.path.join(REF, "tube.in"))
    assert cfg["alpha_Ewd"] == 0.44 and cfg["eps_Ewd"] == 1e-3 and cfg["PBspln_Ewd"] == 8
    assert cfg["nCellTypes"] == 1 and cfg["viscRat"] == [1.0] and cfg["refRad"] == [1.0] and cfg["Deflate"] is False
    assert cfg["Nt"] == 10000 and cfg["Ts"] == 0.0008 and cfg["epsDist"] == 0.02 and cfg["rigidsep"] is False
    assert cfg["restart_file"] == "D/restart.LATEST.dat"
    x, e2v = cases.read_wall_mesh(os.path.join(REF, "new_cyl_D6_L13_33.e"))
    assert x.shape == (3, 1328) and e2v.shape == (3, 2404)
    assert e2v.min() == 1 and e2v.max() == 1328
    # a closed tube surface: every edge belongs to exactly two triangles (the end rings are duplicated vertices)
    r = np.hypot(x[0], x[1])
    assert np.allclose(r, r[0], rtol=1e-6) and abs(x[2].max() - x[2].min() - 13.33) < 1e-2


def test_minicase_configuration_and_oracle_operator(oracle_lib):
    from rbc3d_b200 import cases
    from rbc3d_b200.ewald import SetEwaldPrms
    cfg = cases.read_tube_in(os.path.join(REF, "tube.in"))
    sus, W, vbkg = cases.minicase(os.path.join(REF, "new_cyl_D6_L13_33.e"))
    assert np.allclose(sus.Lb, [10.5, 10.5, 8.0], atol=2e-3)
    orc = oracle_lib.Oracle(sus.Lb, alpha=cfg["alpha_Ewd"], eps=cfg["eps_Ewd"], P=cfg["PBspln_Ewd"])
    assert abs(orc.rc - 1.1986) < 1e-4 and orc.Nb == [48, 48, 36]
    rc, Nb = SetEwaldPrms(sus.Lb, cfg["alpha_Ewd"], cfg["eps_Ewd"], cfg["PBspln_Ewd"])          # product-side ModConf
    assert abs(rc - orc.rc) < 1e-15 and Nb == orc.Nb
    assert sus.ncell == 2 and W.NV == 1328 and W.NE == 2404
    # cells sit inside the tube: distance of every cell point from the axis < tube radius
    d = np.hypot(sus.x[0] - 0.5 * sus.Lb[0], sus.x[1] - 0.5 * sus.Lb[1])
    assert d.max() < 5.0 - 0.5
    assert abs(W.area.sum() - 2 * np.pi * 5.0 * 8.0) / (2 * np.pi * 5.0 * 8.0) < 0.02           # triangulated cylinder
    # operator #3 of the no-slip solve on the real mesh: the single layer of a constant traction on a closed periodic
    # tube plus the cells' single layer, evaluated at the wall vertices -- finite and of the right size
    orc.set_cells(sus)
    orc.set_walls(W)
    W.f[2] = 1.0
    orc.set_wall_traction(W.f)
    orc.prepare_sing_int_on_walls()                      # PrepareSingIntOnWall (TimeInt_Init, ModTimeInt.F90:87)
    rowptr, col, val = orc.wall_matrix(0)
    assert len(rowptr) == 1328 + 1 and np.all(np.diff(rowptr) > 0) and np.all(np.isfinite(val))
    tl = orc.wall_targets()
    v = orc.apply(1.0 / (4 * np.pi), 0.0, tl, cells=True, walls=True)
    assert v.shape == (3, 1328) and np.all(np.isfinite(v)) and np.abs(v).max() > 0
    # by symmetry of the (unrotated) configuration about the tube axis only weakly broken by the two cells, the axial
    # component dominates
    assert np.abs(v[2]).mean() > 5 * np.abs(v[0]).mean()


CAROTID = "/root/reference/examples/carotid_web/Input"


@pytest.mark.skipif(not os.path.isdir(CAROTID), reason="reference tree not mounted")
def test_carotid_web_walls_on_the_oracle(oracle_lib):
    """BASELINE.json configs[4] (wall-dominated operator): the two real wall meshes through the wall matvec of the
    no-slip solve (operator #4) on the oracle."""
    from rbc3d_b200 import cases, noslip
    W, Lb = cases.carotid_web_walls(CAROTID)
    assert list(W.nvert) == [14550, 2903] and list(W.nele) == [28948, 5682]
    assert np.allclose(Lb, [10.5, 10.5, 30.0], atol=2e-3)
    orc = oracle_lib.Oracle(Lb)
    assert orc.Nb == [48, 48, 136] and orc.Nc == [8, 8, 25]
    orc.set_walls(W, ncell=0)
    orc.prepare_sing_int_on_walls()
    # the vessel is periodic in z: its end rings are duplicated vertices; the web has none
    vo = W.voff()
    v2v0 = noslip.wall_build_v2v(W.x[:, vo[0]:vo[1]], Lb)
    v2v1 = noslip.wall_build_v2v(W.x[:, vo[1]:vo[2]], Lb)
    assert (v2v0 > 0).sum() > 50 and (v2v1 > 0).sum() == 0
    # self-interaction matrix: every vertex row is populated; SingIntOnWall = matrix times traction
    rng = np.random.default_rng(4)
    W.f = rng.normal(size=W.f.shape)
    orc.set_wall_traction(W.f)
    for w in range(2):
        rowptr, col, val = orc.wall_matrix(w)
        assert np.all(np.diff(rowptr) > 0) and np.all(np.isfinite(val))
        f = W.f[:, vo[w]:vo[w + 1]]
        v = orc.sing_int_on_wall(0.3, w)
        rows = np.repeat(np.arange(W.nvert[w]), np.diff(rowptr))
        ref = np.zeros_like(v)
        np.add.at(ref.T, rows, np.einsum("bij,bj->bi", val, f.T[col]))
        assert np.abs(v - 0.3 * ref).max() < 1e-12 * np.abs(ref).max()
    # a few GMRES iterations of the wall solve on a uniform slip reduce the residual monotonically
    W.f[:] = 0.0
    from oracle import harness
    rv, mv, st = harness.noslip_backend(orc, [0.0, 0.0, 8.0], cells=False)
    s = noslip.WallNoSlipSolver(W, Lb, rv, mv, st)
    assert s.dof == 3 * (W.NV - int((v2v0 > 0).sum()))
    _, niter, hist, _ = s.solve(rtol=1e-3, maxit=8)
    assert niter == 8 and np.all(np.diff(hist) < 0) and hist[-1] < 0.5 * hist[0]


def test_every_shipped_tube_in_and_wall_mesh_reads():
    """all example inputs of the reference parse: tube.in heads (alpha = 0.44, eps = 1e-3, P = 8 everywhere)
    and every Tri3 Exodus wall mesh under examples/ and sample_files/."""
    import glob
    from rbc3d_b200 import cases
    tubes = sorted(glob.glob("/root/reference/examples/*/Input/tube.in"))
    assert len(tubes) >= 5
    for t in tubes:
        cfg = cases.read_tube_in(t)
        assert (cfg["alpha_Ewd"], cfg["eps_Ewd"], cfg["PBspln_Ewd"]) == (0.44, 1e-3, 8), t
        assert cfg["nCellTypes"] == len(cfg["viscRat"]) >= 1 and all(v == 1.0 for v in cfg["viscRat"]), t
    meshes = sorted(set(glob.glob("/root/reference/examples/*/Input/*.e") + glob.glob("/root/reference/sample_files/*/*/*.e")))
    assert len(meshes) >= 10
    for m in meshes:
        x, e2v = cases.read_wall_mesh(m)
        assert x.shape[0] == 3 and e2v.shape[0] == 3 and e2v.min() == 1 and e2v.max() == x.shape[1], m
        xe = x[:, e2v - 1]
        area = 0.5 * np.linalg.norm(np.cross((xe[:, 1] - xe[:, 0]).T, (xe[:, 2] - xe[:, 0]).T), axis=1)
        assert np.all(area > 0), m                                       # no degenerate triangles


def test_tube_in_list_directed_repeat_counts_and_slash(tmp_path):
    """Fortran list-directed input as ReadConfig's READ(unit, *) accepts it: 'r*c' repeat counts, ',' separators, a '/'
    ending a record's data, D exponents; a '/' inside a quoted file name is data."""
    from rbc3d_b200 import cases
    p = tmp_path / "tube.in"
    p.write_text("0.44 ! alpha\n1d-3\n8\n2\n2*1.0 / rest of the record is not read\n 2.82, 2.9\n.false.\n0.\n0.\n8.\n100\n"
                 "0.0008\n1\n1\n1\n1\n1\n100\n'D/restart.dat'\n0.03\n10.\n4.\n.false.\n0.\n")
    cfg = cases.read_tube_in(str(p))
    assert cfg["viscRat"] == [1.0, 1.0] and cfg["refRad"] == [2.82, 2.9]
    assert cfg["eps_Ewd"] == 1e-3 and cfg["restart_file"] == "D/restart.dat" and cfg["Nt"] == 100
