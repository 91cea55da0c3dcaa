"""TEST INFRASTRUCTURE, NOT PRODUCT CODE: the oracle as the operator behind the harness callers of rbc3d_b200 (the wall
no-slip solve, rbc3d_b200/noslip.py; the boundary-integral work of an mtube time step, rbc3d_b200/mtube.py), so that the
same caller code that drives the CUDA library through the C ABI can be checked against, and timed beside, the CPU
restatement.  Imported only by tests/ and by the cpu_baseline / --impl reference legs of bench.py."""
from __future__ import annotations

import numpy as np

from rbc3d_b200.mtube import C1, VBKG
from rbc3d_b200.noslip import C1_WALL


def noslip_backend(orc, vbkg, cells: bool = True, active=None, collect=None):
    """(residual_vel, wall_matvec, set_traction) on the CPU oracle (tests, cpu_baseline).  Needs orc.set_cells /
    set_walls / prepare_sing_int_on_walls done.  Several ranks: ``active`` = this rank's flags of the wall target list
    (SetActiveFlag), ``collect`` = TargetList_CollectArray (sum over ranks); the background velocity is added after the
    sum, as in Compute_Wall_Residual_Vel (ModNoSlip.F90:186-191)."""
    tl = orc.wall_targets(active)
    vb = np.asarray(vbkg, dtype=float)[:, None]
    collect = collect or (lambda v: v)

    def residual_vel():
        return collect(orc.apply(C1_WALL, C1_WALL, tl, cells=cells, walls=True)) + vb

    def wall_matvec(_f):
        return collect(orc.apply(C1_WALL, 0.0, tl, cells=False, walls=True))

    return residual_vel, wall_matvec, orc.set_wall_traction


class OracleStep:
    """One step's operators on the CPU oracle (bench.py cpu_baseline / --impl reference, tests)."""

    def __init__(self, orc, sus, W, vbkg=VBKG):
        self.orc, self.sus, self.W, self.vbkg = orc, sus, W, np.asarray(vbkg, dtype=float)
        orc.set_cells(sus)
        orc.set_walls(W)
        orc.prepare_sing_int_on_walls()                      # TimeInt_Init

    def update_geometry(self):
        self.orc.set_cells(self.sus)                         # SourceList_UpdateCoord + UpdateDensity, TargetList_Update
        self.orc.set_wall_traction(self.W.f)

    def compute_rhs(self):
        sus = self.sus
        v = self.orc.apply(C1, 0.0, self.orc.cell_targets(), cells=True, walls=True)
        A = np.repeat(sus.Acoef, sus.nlat * sus.nlon)
        return v + 2.0 * self.vbkg[:, None] / A[None, :]

    def noslip_backend(self):
        return noslip_backend(self.orc, self.vbkg)
