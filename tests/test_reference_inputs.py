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
    assert cfg["nCellTypes"] == 1 and cfg["viscRat"] == [1.0] and cfg["Deflate"] is False
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
