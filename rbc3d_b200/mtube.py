"""The boundary-integral work of one ``TimeInt_Euler`` step of examples/minicase/mtube (ModTimeInt.F90:108-176), composed
from the boundary's entry points -- harness code used by bench.py (``mtube`` block: BASELINE.json "mtube timesteps/s")
and by the tests; the same composition runs on the CUDA library (C ABI, ``LibraryStep``) and, for checking and for the
CPU baseline, on the oracle (``oracle/harness.py``: ``OracleStep``).

    Compute_Rbc_Vel   SourceList_UpdateCoord / UpdateDensity(f) on the moved cells         ModTimeInt.F90:127, ModVelSolver.F90:44-93
                      Compute_Rhs = operator #1: c1 = 1/(4 pi), c2 = 0, cells + walls -> cell points, + 2 vBkg / A   :455-515
                      viscRat = 1 in every shipped tube.in: sol = rhs, no cell GMRES                                   :104-105
    NoSlipWall        operator #3 (rhs), operator #4 x wall-GMRES iterations, operator #3 (monitor)                  ModNoSlip.F90:44-149

Membrane forces, spherical-harmonic filtering, volume constraint and repulsion are outside SURVEY.md section 8 and are
not part of the timed step; the traction ``f`` of the cells is a band-limited stand-in (synth.make_suspension).
``PrepareSingIntOnWall`` runs once at start-up (TimeInt_Init, ModTimeInt.F90:87), not per step.
"""
from __future__ import annotations

import os
import time

import numpy as np

from . import noslip, sphere, synth

C1 = 1.0 / (4.0 * np.pi)
VBKG = (0.0, 0.0, 8.0)                                       # examples/minicase/Input/tube.in via mtube.F90


def minicase_like(nlat0: int = 12, dealias: int = 3, seed: int = 161269, ntheta: int = 48, nz: int = 26):
    """examples/minicase with a generated tube mesh (the Exodus file is not on the GPU box): box 10.5 x 10.5 x 8, tube
    radius 5 (minit.F90:39-70), two unrotated biconcave cells at the init program's positions (:80-96), lambda = 1.
    48 x 26 gives 1296 vertices / 2496 triangles against the file's 1328 / 2404."""
    Lb = np.array([10.5, 10.5, 8.0])
    centers = np.array([[-0.5, -0.5, 4.0], [0.5, 0.5, 1.0]]) + np.array([0.5 * Lb[0], 0.5 * Lb[1], 0.0])
    sus = synth.make_suspension(1, nlat0=nlat0, dealias=dealias, seed=seed, L=1.0, centers=centers, rotate=False,
                                visc_ratio=1.0)
    sus.Lb = Lb
    W = synth.make_walls(Lb, [dict(radius=5.0, ntheta=ntheta, nz=nz)])
    W.f[:] = 0.0                                             # minit.F90 writes zero wall tractions into the restart file
    return sus, W


GOLDEN_SICKLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "ref_sickle_cell.npz")


def import_read_rbc(x_file: np.ndarray, xc) -> np.ndarray:
    """ImportReadRBC (ModIO.F90:668-681): the exported cell recentred by (5.25, 5.25, 5/7), then moved to xc.
    x_file (3, nlon, nlat) as parsed from SickleCell.dat (scripts/make_golden_sickle.py)."""
    off = np.array([5.25, 5.25, 5.0 / 7.0])
    return x_file - off[:, None, None] + np.asarray(xc, dtype=float)[:, None, None]


def case_like(nrbc: int = 8, sickles: bool = True, seed: int = 161269, ntheta: int = 48, nz: int = 36,
              visc_ratio: float = 1.0, sickle_x: np.ndarray | None = None):
    """examples/case (initcond.F90:42-110) and examples/case_sickles (sickle_initcond.F90:50-110) with a generated tube
    mesh: tube radius 5, length nrbc / 0.7, box 10.5 x 10.5 x length, the cells on the axis at z = (iz - 1/2) length /
    nrbc, all biconcave (case) or every second one the imported sickle cell (case_sickles; shape from the committed
    input file rbc3d_b200/data/ref_sickle_cell.npz = the reference's SickleCell.dat).  BASELINE.json configs[1] / configs[3]."""
    length = nrbc / 0.7
    Lb = np.array([10.5, 10.5, length])
    spacing = length / nrbc
    th, phi, _ = sphere.gauss_grid(36, 72)
    xb, _, _ = sphere.biconcave_unit(th, phi, 1.0)
    if sickles and sickle_x is None:
        sickle_x = np.load(GOLDEN_SICKLE)["x"]
    xs = []
    for iz in range(1, nrbc + 1):
        xc = np.array([0.5 * Lb[0], 0.5 * Lb[1], spacing * (iz - 0.5)])          # after Recenter_Cells_and_Walls
        if sickles and iz % 2 == 0:
            xs.append(import_read_rbc(sickle_x, xc))
        else:
            xs.append(xb + xc[:, None, None])
    sus = synth.suspension_from_shapes(np.stack(xs), Lb, nlat0=12, dealias=3, visc_ratio=visc_ratio, seed=seed)
    W = synth.make_walls(Lb, [dict(radius=5.0, ntheta=ntheta, nz=nz)])
    W.f[:] = 0.0
    return sus, W


TS = 0.0008                                                  # examples/minicase/Input/tube.in: Ts


def advect_rigid(sus, v_cells: np.ndarray, Ts: float = TS) -> None:
    """Stand-in for ``rbc%x = rbc%x + Ts*rbc%v`` (ModTimeInt.F90:136-141) + ReboxRbcs that needs no membrane solver: every
    cell is translated by Ts times its mean surface velocity, so x and spline(x) move and every other geometric field
    (a3, detJ, their splines, the densities) is unchanged.  Gives each time step of the harness a NEW geometry, which
    is what makes the cell lists, the geometry caches and the wall solve's right-hand side change from step to step."""
    npc = sus.nlat * sus.nlon
    for c in range(sus.ncell):
        d = Ts * v_cells[:, c * npc:(c + 1) * npc].mean(axis=1)
        ctr = sus.centers[c] + d
        wrap = -np.floor(ctr / sus.Lb) * sus.Lb              # ReboxRbcs: keep the centre inside the box
        d = d + wrap
        sus.centers[c] = sus.centers[c] + d
        sus.x[:, c * npc:(c + 1) * npc] += d[:, None]
        sus.spx[c, 0] += d[:, None, None]                    # spline values u; derivatives u1, u2, u12 do not change


class LibraryStep:
    """The same on the CUDA library through the C ABI (rbc3d_b200.ewald.EwaldOperator)."""

    def __init__(self, op, sus, W, vbkg=VBKG, device_noslip=False):
        self.op, self.sus, self.W, self.vbkg = op, sus, W, np.asarray(vbkg, dtype=float)
        self.device_noslip = device_noslip       # NoSlipWall resident on the GPU (rbc3d_noslip_solve)
        op.set_suspension(sus)
        op.set_walls(W)
        op.PrepareSingIntOnWall()

    def update_geometry(self):
        sus, op = self.sus, self.op
        op.SourceList_UpdateCoord(sus.x, sus.a3, sus.Acoef, sus.Bcoef, sus.area, sus.meshSize, sus.spx, sus.spa3, sus.spdetj)
        op.SourceList_UpdateDensity(sus.weighted(sus.f), sus.weighted(sus.g), sus.spF, sus.spG)
        op.set_wall_traction(self.W.f)

    def compute_rhs(self):
        from .capi import TL_CELLS
        sus = self.sus
        v = self.op.apply(C1, 0.0, TL_CELLS, cells=True, walls=True)
        A = np.repeat(sus.Acoef, sus.nlat * sus.nlon)
        return v + 2.0 * self.vbkg[:, None] / A[None, :]

    def noslip_backend(self):
        return noslip.library_backend(self.op, self.vbkg)


def bi_timestep(step, rtol: float = 1e-3, maxit: int = 60, advect: bool = False):
    """Run the step's boundary-integral work once (advect: then move the cells for the next step, untimed).  -> dict(v_cells, f_wall, wall_iterations, history, slip, seconds{})."""
    t = {}
    t0 = time.perf_counter()
    step.update_geometry()
    t["geometry"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    v = step.compute_rhs()
    t["rhs"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    if getattr(step, "device_noslip", False):
        f, niter, hist, slip = noslip.solve_on_device(step.op, step.W, step.sus.Lb, step.vbkg, rtol=rtol, maxit=maxit)
        nmatvec = niter + (niter - 1) // 30      # one operator #4 per iteration + the true residual at every restart
    else:
        s = noslip.WallNoSlipSolver(step.W, step.sus.Lb, *step.noslip_backend())
        f, niter, hist, slip = s.solve(rtol=rtol, maxit=maxit)
        nmatvec = s.nmatvec
    t["noslip"] = time.perf_counter() - t0
    t["total"] = t["geometry"] + t["rhs"] + t["noslip"]
    if advect:
        advect_rigid(step.sus, v)                            # untimed: the caller's position update
    return {"v_cells": v, "f_wall": f, "wall_iterations": niter, "history": hist, "slip": slip, "seconds": t,
            "operator_applications": 1 + 2 + nmatvec}
