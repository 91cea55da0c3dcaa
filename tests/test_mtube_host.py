"""The boundary-integral work of one mtube time step (rbc3d_b200/mtube.py) on the CPU oracle: BASELINE.json configs[0]
(examples/minicase) with a generated tube mesh, and -- where /root/reference is mounted -- with the reference's own
Exodus mesh.  PARITY UNPINNED (no recorded output of the reference exists); the checks are physical: the no-slip
solve cancels the wall slip, cells in a tube flow are carried along the axis, tractions oppose the flow."""
import os

import numpy as np
import pytest

from oracle import harness
from rbc3d_b200 import mtube

MESH = "/root/reference/examples/minicase/Input/new_cyl_D6_L13_33.e"


def check_step(r, sus, first):
    assert 0 < r["wall_iterations"] <= 60
    assert r["history"][-1] < 1e-3 * r["history"][0]
    assert r["operator_applications"] >= 3 + r["wall_iterations"]
    assert r["v_cells"].shape == (3, sus.npoint) and np.all(np.isfinite(r["v_cells"]))
    if first:
        # zero wall tractions: the cells see the background flow (2 vBkg / A with A = 2) plus their own single layer
        assert abs(r["v_cells"][2].mean() - 8.0) < 1.0


def test_two_steps_generated_mesh(oracle_lib):
    sus, W = mtube.minicase_like(nlat0=6)
    step = harness.OracleStep(oracle_lib.Oracle(sus.Lb), sus, W)
    r1 = mtube.bi_timestep(step)
    check_step(r1, sus, True)
    slip0 = r1["history"][0]
    r2 = mtube.bi_timestep(step)                       # tractions carried over: the start residual is the end residual
    check_step(r2, sus, False)
    assert abs(r2["history"][0] - r1["history"][-1]) < 1e-6 * slip0
    # vBkg is the mean velocity of the periodic box; with the wall holding the fluid at rest the profile becomes
    # Poiseuille-like, so the cells near the axis move faster than the mean (about 2x for a tube) and the wall
    # traction opposes the flow
    assert 1.5 * 8.0 < r2["v_cells"][2].mean() < 3.0 * 8.0
    assert r2["f_wall"][2].mean() < 0


@pytest.mark.skipif(not os.path.exists(MESH), reason="reference tree not mounted")
def test_one_step_reference_mesh(oracle_lib):
    from rbc3d_b200 import cases
    sus, W, vbkg = cases.minicase(MESH, nlat0=6)
    step = harness.OracleStep(oracle_lib.Oracle(sus.Lb), sus, W, vbkg)
    r = mtube.bi_timestep(step)
    check_step(r, sus, True)
    # 1328 vertices / 2404 triangles: an open cylinder has 2 V - F = 252 boundary vertices = two end rings of 126 that
    # coincide modulo the period
    from rbc3d_b200 import noslip
    v2v = noslip.wall_build_v2v(W.x, sus.Lb)
    ndup = int((v2v > 0).sum())
    assert ndup == 126 and r["f_wall"].shape == (3, 1328)
    assert np.array_equal(r["f_wall"][:, v2v > 0], r["f_wall"][:, v2v[v2v > 0] - 1])


def test_rigid_advection_keeps_the_splines_consistent():
    from rbc3d_b200 import synth
    sus, _ = mtube.minicase_like(nlat0=6)
    rng = np.random.default_rng(2)
    v = rng.normal(size=(3, sus.npoint)) + np.array([0.0, 0.0, 9000.0])[:, None]      # far enough to leave the box in z
    x0, c0 = sus.x.copy(), sus.centers.copy()
    mtube.advect_rigid(sus, v)
    d = sus.centers - c0
    assert np.all(sus.centers >= 0) and np.all(sus.centers < sus.Lb)                  # ReboxRbcs
    npc = sus.nlat * sus.nlon
    for c in range(sus.ncell):
        assert np.allclose(sus.x[:, c * npc:(c + 1) * npc] - x0[:, c * npc:(c + 1) * npc], d[c][:, None], atol=1e-12)
    moved = sus.spx.copy()
    synth.build_splines(sus, sus._builder, which=("x",))                              # Rbc_BuildSurfaceSource(xFlag) afresh
    assert np.abs(moved - sus.spx).max() < 1e-11
