"""NoSlipWall around the drop-in boundary (harness code, not product): the first-kind equation for the wall tractions
that every time step solves after the cell velocities (ModTimeInt.F90:133, ModNoSlip.F90:44-149).

    rhs  = -Compute_Wall_Residual_Vel          operator #3: c1 = c2 = 1/(4 pi), cells + walls -> wall vertices, + vBkg
    A df = MyMatMult(df)                       operator #4: c1 = 1/(4 pi), walls -> wall vertices   (ModNoSlip.F90:255-308)
    KSPGMRES, PCNONE, rtol = eps_Ewd, at most 60 iterations, zero initial guess                    (:70-87, 117)
    wall%f = f0 + df ; residual velocity monitored with a second operator #3                       (:131-146)

The unknowns are the tractions of the vertices without periodic duplicates: ``Wall_Build_V2V`` (ModWall.F90:64-113)
finds the duplicated boundary rings, ``indxVertGlb`` (ModData.F90:50-63) numbers the independent vertices and
``AssembleArray`` (ModNoSlip.F90:362-384) moves between the two numberings -- duplicates are overwritten in vertex
order going to 1-D, so the last duplicate wins, exactly as the reference's loop does.

The solver is written against callables so that the same code runs on the CUDA library through the C ABI
(``library_backend``; GPU tests, bench.py) and, for the tests and the CPU baseline, on the oracle (``oracle/harness.py``): ``residual_vel() -> v (3, NV)`` and ``wall_matvec(f (3, NV)) -> v (3, NV)``.
"""
from __future__ import annotations

import numpy as np

from .gmres import gmres

C1_WALL = 1.0 / (4.0 * np.pi)            # ModNoSlip.F90:172-173, 281


def wall_build_v2v(x: np.ndarray, Lb) -> np.ndarray:
    """Wall_Build_V2V for one wall: x (3, nvert) -> v2v (nvert,) int32, 1-based master vertex or 0."""
    Lb = np.asarray(Lb, dtype=float)
    iLb = 1.0 / Lb
    eps = Lb.min() * 1.0e-5                                        # -freal-4-real-8: the literal 1.E-5 is a double
    xmin, xmax = x.min(axis=1), x.max(axis=1)
    bd = np.nonzero((np.abs(x - xmin[:, None]).min(axis=0) < eps) | (np.abs(x - xmax[:, None]).min(axis=0) < eps))[0]
    v2v = np.zeros(x.shape[1], dtype=np.int32)
    for p1 in range(len(bd) - 1):
        i1 = bd[p1]
        if v2v[i1] > 0:
            continue
        rest = bd[p1 + 1:]
        xx = x[:, rest] - x[:, [i1]]
        xx = xx - np.rint(xx * iLb[:, None]) * Lb[:, None]
        hit = rest[np.abs(xx).max(axis=0) < eps]
        v2v[hit] = i1 + 1
    return v2v


def indx_vert_glb(v2v_per_wall) -> tuple[np.ndarray, int]:
    """GlobData_Init, ModData.F90:50-63: 1-based global number of every wall vertex, duplicates sharing their
    master's number.  -> (indx over all walls back to back, number of independent vertices)."""
    out, p = [], 0
    for v2v in v2v_per_wall:
        idx = np.zeros(len(v2v), dtype=np.int64)
        for i, m in enumerate(v2v):
            if m == 0:
                p += 1
                idx[i] = p
            else:
                idx[i] = idx[m - 1]
        out.append(idx)
    return np.concatenate(out), p


class WallNoSlipSolver:
    """NoSlipWall.  W: rbc3d_b200.synth.Walls; Lb: box; residual_vel / wall_matvec: see the module docstring.
    ``set_traction(f)`` is called with the traction the operators must see (wall%f = ...)."""

    def __init__(self, W, Lb, residual_vel, wall_matvec, set_traction):
        self.W, self.Lb = W, np.asarray(Lb, dtype=float)
        vo = W.voff()
        self.v2v = [wall_build_v2v(W.x[:, vo[w]:vo[w + 1]], self.Lb) for w in range(W.nwall)]
        self.indx, self.nindep = indx_vert_glb(self.v2v)
        self.dof = 3 * self.nindep                                # ModNoSlip.F90:62-67
        self.residual_vel, self.wall_matvec, self.set_traction = residual_vel, wall_matvec, set_traction
        self.nmatvec = 0

    # AssembleArray(u, u1D, +1 / -1): u1D(3 p - 2 : 3 p) = u(ivert, :)
    def to_1d(self, u: np.ndarray) -> np.ndarray:
        u1 = np.zeros(self.dof)
        idx = self.indx - 1
        for d in range(3):
            u1[3 * idx + d] = u[d]                                # numpy keeps the last write of a repeated index
        return u1

    def from_1d(self, u1: np.ndarray) -> np.ndarray:
        idx = self.indx - 1
        return np.ascontiguousarray(np.stack([u1[3 * idx + d] for d in range(3)]))

    def matmult(self, u1: np.ndarray) -> np.ndarray:
        f = self.from_1d(u1)
        self.set_traction(f)                                      # wall%f = f, ModNoSlip.F90:273-278
        self.nmatvec += 1
        return self.to_1d(self.wall_matvec(f))

    def solve(self, rtol: float = 1e-3, maxit: int = 60):
        """-> (f_new (3, NV), niter, residual history, residual velocity after the update (3, NV))."""
        W = self.W
        f0 = np.array(W.f, dtype=float)
        self.set_traction(f0)
        rhs = self.to_1d(-self.residual_vel())
        df1, niter, hist = gmres(self.matmult, rhs, x0=None, rtol=rtol, maxit=maxit)
        f_new = f0 + self.from_1d(df1)
        self.set_traction(f_new)
        W.f = f_new
        return f_new, niter, hist, self.residual_vel()


def solve_on_device(op, W, Lb, vbkg, cells: bool = True, rtol: float = 1e-3, maxit: int = 60):
    """NoSlipWall with tractions, Krylov vectors and both operators resident on the GPU (rbc3d_noslip_solve): the host
    only supplies indxVertGlb.  Same return value as WallNoSlipSolver.solve; W.f is updated."""
    vo = W.voff()
    v2v = [wall_build_v2v(W.x[:, vo[w]:vo[w + 1]], np.asarray(Lb, dtype=float)) for w in range(W.nwall)]
    indx, nindep = indx_vert_glb(v2v)
    f_new, niter, hist, slip = op.noslip_solve(W.f, indx, nindep, vbkg, cells=cells, rtol=rtol, maxit=maxit)
    W.f = f_new
    return f_new, niter, list(hist), slip


def library_backend(op, vbkg, cells: bool = True, collect: bool = False):
    """The same three callables on the CUDA library through the C ABI (rbc3d_b200.ewald.EwaldOperator with
    set_suspension / set_walls / PrepareSingIntOnWall done).  ``collect``: several ranks -- rbc3d_apply_collect sums the
    rows over the ranks on the devices (the contexts were given their ``active`` flags in set_walls)."""
    from .capi import TL_WALLS
    vb = np.asarray(vbkg, dtype=float)[:, None]
    apply = op.apply_collect if collect else op.apply

    def residual_vel():
        return apply(C1_WALL, C1_WALL, TL_WALLS, cells=cells, walls=True) + vb

    def wall_matvec(_f):
        return apply(C1_WALL, 0.0, TL_WALLS, cells=False, walls=True)

    return residual_vel, wall_matvec, op.set_wall_traction
