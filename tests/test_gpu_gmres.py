"""GMRES parity (BASELINE.json north_star: "full GMRES solves must converge to the same residual in the same or fewer
iterations"): Solve_RBC_Vel restated on the host (rbc3d_b200/gmres.py: PETSc-default GMRES(30), Glob_Sph_Trans packing)
driven once by the GPU library through the C ABI and once by the CPU oracle, on the same 8-cell suspension."""
import copy

import numpy as np
import pytest

from tests.util import C1_RHS, C2_MATVEC, rel_l2, small_suspension

pytestmark = pytest.mark.gpu


def test_cell_gmres_solve_gpu_vs_oracle():
    from oracle import oracle
    from rbc3d_b200 import gmres, synth
    from rbc3d_b200.ewald import EwaldOperator
    sus = small_suspension(2)
    op = EwaldOperator(sus.Lb)
    op.set_suspension(sus)
    op.enable_device_splines(sus.nlat0)          # Rbc_BuildSurfaceSource(gFlag) on the device, as bench.py does

    def gpu_sl(fw):
        op.SourceList_UpdateDensity(f=fw)
        return op.apply(C1_RHS, 0.0)

    def gpu_dl(gw, g_raw):
        op.SourceList_UpdateDensity(g=gw)
        return op.apply(0.0, C2_MATVEC)

    osus = copy.copy(sus)
    orc = oracle.Oracle(sus.Lb)

    def cpu_sl(fw):
        orc.set_cells(osus)
        return orc.apply_cells(C1_RHS, 0.0, orc.cell_targets())

    def cpu_dl(gw, g_raw):
        osus.g = np.ascontiguousarray(g_raw)
        synth.build_splines(osus, sus._builder, which=("G",))
        orc.set_cells(osus)
        return orc.apply_cells(0.0, C2_MATVEC, orc.cell_targets())

    sg = gmres.CellVelocitySolver(sus, gpu_sl, gpu_dl)
    sc = gmres.CellVelocitySolver(sus, cpu_sl, cpu_dl)
    rhs_g, rhs_c = sg.compute_rhs(), sc.compute_rhs()
    assert rel_l2(rhs_g, rhs_c) < 1e-10
    sol_g, v_g, it_g, h_g = sg.solve(rhs=rhs_g, rtol=1e-11)
    sol_c, v_c, it_c, h_c = sc.solve(rhs=rhs_c, rtol=1e-11)
    print(f"GMRES: gpu {it_g} its, residual {h_g[-1]:.3e}; oracle {it_c} its, residual {h_c[-1]:.3e}; "
          f"|v_gpu - v_cpu| / |v_cpu| = {rel_l2(v_g, v_c):.2e}")
    assert 0 < it_g <= it_c < 200
    assert h_g[-1] < 1e-11 * np.linalg.norm(rhs_g)
    k = min(len(h_g), len(h_c))
    assert np.allclose(h_g[:k], h_c[:k], rtol=1e-6, atol=1e-9 * h_c[0])
    assert rel_l2(v_g, v_c) < 1e-8
    op.close()


def test_device_resident_solver_matches_host_harness():
    """MyMatMult and the GMRES solve with SH transforms and Krylov vectors on the device
    (rbc3d_solver_*) against the host harness driving the same GPU operator through host buffers."""
    from rbc3d_b200 import gmres
    from rbc3d_b200.ewald import EwaldOperator
    sus = small_suspension(2)
    op = EwaldOperator(sus.Lb)
    op.set_mesh(sus.ncell, sus.nlat, sus.nlon, sus.th, sus.phi, sus.w)
    op.enable_device_splines(sus.nlat0)
    op.SourceList_UpdateCoord_mesh(sus.x, sus.a3, sus.detj, sus.Acoef, sus.Bcoef, sus.area, sus.meshSize)
    op.SourceList_UpdateDensity(f=sus.weighted(sus.f), g=sus.weighted(sus.g))

    def gpu_sl(fw):
        op.SourceList_UpdateDensity(f=fw)
        return op.apply(C1_RHS, 0.0)

    def gpu_dl(gw, g_raw):
        op.SourceList_UpdateDensity(g=gw)
        return op.apply(0.0, C2_MATVEC)

    host = gmres.CellVelocitySolver(sus, gpu_sl, gpu_dl)
    rhs = host.compute_rhs()
    op.solver_setup(sus.nlat0, sus.detj)
    assert op.solver_dof == host.T.dof
    rng = np.random.default_rng(4)
    u = rng.uniform(-1, 1, host.T.dof)
    assert rel_l2(op.solver_matmult(u), host.matmult(u)) < 1e-11          # same transforms, same operator
    op.SourceList_UpdateDensity(f=sus.weighted(sus.f))
    assert rel_l2(op.solver_rhs((1.0, 0.0, 0.0)), rhs) < 1e-11              # Compute_Rhs on the device
    sol_h, v_h, it_h, hist_h = host.solve(rhs=rhs, rtol=1e-11)
    sol_d, it_d, hist_d = op.solver_gmres(rhs, rtol=1e-11)
    assert rel_l2(op.solver_velocity(sol_d), v_h) < 1e-8                     # Glob_Sph_Trans(v, sol, FOUR_TO_PHYS)
    print(f"device-resident GMRES: {it_d} its, residual {hist_d[-1]:.3e}; host harness: {it_h} its, {hist_h[-1]:.3e}; "
          f"|sol_d - sol_h| / |sol_h| = {rel_l2(sol_d, sol_h):.2e}")
    assert it_d == it_h
    assert np.allclose(hist_d, hist_h, rtol=1e-6, atol=1e-9 * hist_h[0])
    assert rel_l2(sol_d, sol_h) < 1e-8
    # a second solve from the converged solution needs no iteration (KSPSetInitialGuessNonzero)
    _, it2, _ = op.solver_gmres(rhs, x0=sol_d, rtol=1e-10)
    assert it2 == 0
    op.close()
