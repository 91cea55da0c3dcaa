"""The BASELINE.json example configurations on the reference's OWN wall meshes (committed fixtures under
rbc3d_b200/data/meshes/, scripts/make_golden_meshes.py) through the C ABI, against the oracle: operators #1 - #4 of a time
step for examples/minicase, examples/case, examples/case_sickles and examples/carotid_web with its 72
cells and two walls (carotid_initcond.F90:15-54; cell placement from rbc3d_b200/data/carotid_web_cells.npz).

Tolerance: relative L2 <= 1e-10 over all targets (north_star); cell ids bit-exact; wall GMRES with the same or fewer
iterations.  lambda = 5 instead of the examples' 1 so that operator #2 (double layer, coefficient 1 - lambda) is not
identically zero."""
import os

import numpy as np
import pytest

from . import util
from .util import C1_RHS, C2_MATVEC, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _operators(orc, with_matvec=True):
    from rbc3d_b200.capi import TL_CELLS, TL_WALLS
    ops = [("#1 Compute_Rhs", C1_RHS, 0.0, TL_CELLS, "cells", True, True),
           ("#3 Compute_Wall_Residual_Vel", C1_RHS, C1_RHS, TL_WALLS, "walls", True, True),
           ("#4 wall MyMatMult", C1_RHS, 0.0, TL_WALLS, "walls", False, True)]
    if with_matvec:
        ops.insert(1, ("#2 cell MyMatMult", 0.0, C2_MATVEC, TL_CELLS, "cells", True, False))
    return ops


def _check_all(op, orc, sus, W, seed=3):
    rng = np.random.default_rng(seed)
    W.f = rng.normal(size=W.f.shape)
    op.set_suspension(sus)
    op.set_walls(W)
    op.PrepareSingIntOnWall()
    orc.set_cells(sus)
    orc.set_walls(W)
    orc.prepare_sing_int_on_walls()
    assert np.array_equal(op.cell_list()[1], orc.cell_ids(sus.x))                      # bit-exact cell ids
    cnt, sig = op.neighbor_signature()
    rcnt, rsig = orc.neighbor_signature(sus.x, sus.x)
    assert np.array_equal(cnt, rcnt) and np.array_equal(sig, rsig)                     # bit-exact in-range neighbour sets
    errs = {}
    for name, c1, c2, kind, tl_name, cells, walls in _operators(orc):
        tl = orc.cell_targets() if tl_name == "cells" else orc.wall_targets()
        v = op.apply(c1, c2, kind, cells=cells, walls=walls)
        ref = orc.apply(c1, c2, tl, cells=cells, walls=walls)
        errs[name] = rel_l2(v, ref)
        assert np.abs(ref).max() > 0
    print("rel L2 vs oracle:", {k: "%.2e" % e for k, e in errs.items()})
    assert max(errs.values()) < TOL, errs
    return errs


@pytest.mark.parametrize("config", ["minicase", "case", "case_sickles"])
def test_example_configurations_on_the_reference_wall_mesh(oracle_lib, config):
    from rbc3d_b200 import cases
    from rbc3d_b200.ewald import EwaldOperator
    if config == "minicase":
        sus, W, vbkg = cases.minicase(cases.mesh_file("new_cyl_D6_L13_33.e"))
        # lambda = 5 (A = 6, B = -4)
        sus.Acoef[:], sus.Bcoef[:] = 6.0, -4.0
        want_nb = [48, 48, 36]
    else:
        sus, W, vbkg = cases.case(nrbc=8, sickles=(config == "case_sickles"), visc_ratio=5.0)
        want_nb = [48, 48, 52]
    assert W.NV == 1328 and W.NE == 2404
    op = EwaldOperator(sus.Lb)
    orc = oracle_lib.Oracle(sus.Lb)
    assert list(op.Nb) == orc.Nb == want_nb
    _check_all(op, orc, sus, W)
    # the wall no-slip solve around operators #3 / #4 (ModNoSlip.F90:44-149): same or fewer iterations than the oracle
    import copy
    from oracle import harness
    from rbc3d_b200 import noslip
    out = []
    for backend in (harness.noslip_backend(orc, vbkg), noslip.library_backend(op, vbkg)):
        Wc = copy.copy(W)
        Wc.f = np.zeros_like(W.f)
        out.append(noslip.WallNoSlipSolver(Wc, sus.Lb, *backend).solve(rtol=1e-3, maxit=60))
    (f_o, it_o, h_o, _), (f_g, it_g, h_g, _) = out
    assert 0 < it_g <= it_o <= 60
    n = min(len(h_o), len(h_g))
    assert np.allclose(h_g[:n], h_o[:n], rtol=1e-5, atol=1e-9 * h_o[0])
    op.close()


def test_carotid_web_72_cells_and_two_walls(oracle_lib):
    """BASELINE.json configs[4] as the example builds it: carotid.e (14 550 vertices / 28 948 triangles) + web.e (2 903 /
    5 682), 72 cells, box 10.5 x 10.5 x 30, PME grid 48 x 48 x 136, real-space cells 8 x 8 x 25."""
    from rbc3d_b200 import cases
    from rbc3d_b200.ewald import EwaldOperator
    if not os.path.exists(cases.CAROTID_CELLS):
        pytest.skip("rbc3d_b200/data/carotid_web_cells.npz missing (scripts/make_golden_carotid_cells.py)")
    pl = np.load(cases.CAROTID_CELLS)
    sus, W, Lb, vbkg = cases.carotid_web(placement=(pl["centres"], pl["rotations"]), visc_ratio=5.0)
    assert sus.ncell == 72 and sus.npoint == 186624 and list(W.nvert) == [14550, 2903] and list(W.nele) == [28948, 5682]
    op = EwaldOperator(Lb)
    orc = oracle_lib.Oracle(Lb)
    assert list(op.Nb) == orc.Nb == [48, 48, 136] and orc.Nc == [8, 8, 25]
    _check_all(op, orc, sus, W)
    assert op.cell_list_dims() == [8, 8, 25]
    # a few iterations of the wall solve with the cells present: identical residual history
    import copy
    from oracle import harness
    from rbc3d_b200 import noslip
    out = []
    for backend in (harness.noslip_backend(orc, vbkg), noslip.library_backend(op, vbkg)):
        Wc = copy.copy(W)
        Wc.f = np.zeros_like(W.f)
        out.append(noslip.WallNoSlipSolver(Wc, Lb, *backend).solve(rtol=1e-3, maxit=5))
    (_, it_o, h_o, _), (_, it_g, h_g, _) = out
    assert it_g == it_o == 5 and np.allclose(h_g, h_o, rtol=1e-6)
    op.close()
