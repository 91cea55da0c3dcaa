"""Edge cases of the operator on the oracle (CPU): empty and all-inactive target lists, a target on top of a source
point, a suspension of one cell, no cells at all with walls present, zero densities."""
import numpy as np

from rbc3d_b200 import synth
from tests import util
from tests.util import C1_RHS, C2_MATVEC


def test_empty_and_inactive_target_lists(oracle_lib):
    sus = util.small_suspension(2, nlat0=4)
    orc = oracle_lib.Oracle(sus.Lb).set_cells(sus)
    v = orc.apply_cells(C1_RHS, C2_MATVEC, orc.make_targets(np.zeros((3, 0))))
    assert v.shape == (3, 0)
    act = np.zeros(sus.npoint, np.int32)
    v0 = np.full((3, sus.npoint), 3.25)
    v = orc.apply_cells(C1_RHS, C2_MATVEC, orc.cell_targets(active=act), v=v0.copy())
    assert np.array_equal(v, v0)                                   # rows of inactive targets are left untouched


def test_raw_target_on_top_of_a_source_point(oracle_lib):
    """r < 1e-3 sqrt(alpha / pi): both Ewald kernels return zero (ModEwaldFunc.F90:109-113, 164-167), so a raw target
    that coincides with a mesh point stays finite; moving it by 1e-9 changes nothing beyond the smoothness of the rest."""
    sus = util.small_suspension(2, nlat0=4)
    orc = oracle_lib.Oracle(sus.Lb).set_cells(sus)
    x0 = sus.x[:, [17, 300]].copy()
    v_on = orc.apply_cells(C1_RHS, C2_MATVEC, orc.make_targets(x0))
    assert np.all(np.isfinite(v_on))
    v_near = orc.apply_cells(C1_RHS, C2_MATVEC, orc.make_targets(x0 + 1e-9))
    assert np.abs(v_on - v_near).max() < 1e-6 * np.abs(v_on).max()


def test_single_cell_and_zero_density(oracle_lib):
    sus = synth.make_suspension(1, nlat0=4, L=6.0, centers=np.array([[3.0, 3.0, 3.0]]))
    orc = oracle_lib.Oracle(sus.Lb).set_cells(sus)
    v = orc.apply_cells(C1_RHS, C2_MATVEC, orc.cell_targets())
    assert v.shape == (3, sus.npoint) and np.all(np.isfinite(v)) and np.abs(v).max() > 0
    sus.f[:] = 0.0
    sus.g[:] = 0.0
    synth.build_splines(sus, sus._builder, which=("F", "G"))
    orc.set_cells(sus)
    assert not orc.apply_cells(C1_RHS, C2_MATVEC, orc.cell_targets()).any()      # linear operator: zero in, zero out


def test_walls_without_cells_and_zero_traction(oracle_lib):
    Lb = np.array([10.5, 10.5, 8.0])
    W = synth.make_walls(Lb, [dict(radius=4.4, ntheta=16, nz=8)])
    orc = oracle_lib.Oracle(Lb)
    orc.set_walls(W, ncell=0)
    orc.prepare_sing_int_on_walls()
    tl = orc.wall_targets()
    v = orc.apply(C1_RHS, 0.0, tl, cells=False, walls=True)
    assert np.all(np.isfinite(v)) and np.abs(v).max() > 0
    orc.set_wall_traction(np.zeros_like(W.f))
    assert not orc.apply(C1_RHS, 0.0, tl, cells=False, walls=True).any()
    # c1 = 0: nothing is spread, nothing is integrated (flag_sing_lay = |c1| > 1e-10, ModPME.F90:71)
    orc.set_wall_traction(W.f)
    assert not orc.apply(0.0, 0.0, tl, cells=False, walls=True).any()


def test_set_ewald_prms_three_ways(oracle_lib):
    """SetEwaldPrms (ModConf.F90:348-408) restated in the oracle (C), in the library's host arithmetic
    (rbc3d_set_ewald_prms) and here in NumPy from the Fortran; the box sizes of BASELINE.json's configurations."""
    from rbc3d_b200.ewald import SetEwaldPrms
    for Lb, nranks in (((10.5, 10.5, 8.0), 1), ((10.5, 10.5, 8.0), 2), ((10.5, 10.5, 8 / 0.7), 1), ((10.5, 10.5, 30.0), 8),
                       ((57.2589, 57.2589, 57.2589), 8), ((3.0, 2.5, 2.0), 3)):
        Lb = np.array(Lb)
        alpha, eps, P = 0.44, 1e-3, 8
        s = 1.0
        for _ in range(10):
            s = 0.75 * np.sqrt(np.pi) * eps / (s ** 3 + 1.5 * s + 0.75 / s)
            s = np.sqrt(-np.log(s))
        rc = min((Lb / 3.001).min(), np.sqrt(alpha / np.pi) * s)
        Nb = (2 * np.ceil(np.sqrt(-np.log(eps) / (np.pi * alpha)) * Lb)).astype(int)
        Nb[2] = max(Nb[2], nranks * P)
        Nb[2] = int(np.ceil(Nb[2] / nranks)) * nranks
        orc = oracle_lib.Oracle(Lb, alpha, eps, P, nranks=nranks)
        rc_lib, Nb_lib = SetEwaldPrms(Lb, alpha, eps, P, nranks=nranks)
        assert abs(orc.rc - rc) < 1e-15 and abs(rc_lib - rc) < 1e-15
        assert orc.Nb == list(Nb) == list(Nb_lib)
        assert orc.Nc == [max(int(np.floor(L / rc)), 3) for L in Lb]      # HashTable_ComputeNumBlocks
