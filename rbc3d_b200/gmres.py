"""Host-side mirror of the cell velocity solve around the boundary-integral operator (ModVelSolver.F90).

This is harness code for the north-star claim "full GMRES solves converge to the same residual in the same or fewer
iterations": the Krylov solver and the spherical-harmonic packing of the unknowns stay on the host (they are PETSc and
SPHEREPACK in the reference), every operator application goes through the drop-in boundary.

* ``ShTransform`` / ``glob_sph_trans``  -- Glob_Sph_Trans (ModVelSolver.F90:641-719) with ShAnalGau / ShSynthGau
  (ModSphpk.F90:74-105, 240-270) restated.  SPHEREPACK 3.2 is not vendored under /root/reference; the convention used
  here is the documented one of shags/shsgs:  f = sum_n [ a(0,n)/2 Pbar_n^0 + sum_{m>=1} Pbar_n^m (a(m,n) cos m phi -
  b(m,n) sin m phi) ],  a, b = (1/pi) int f Pbar_n^m (cos, -sin)(m phi) dOmega,  int_0^pi Pbar^2 sin = 1.
* ``gmres``  -- KSPGMRES with the defaults the reference leaves untouched: restart 30, classical
  Gram-Schmidt without refinement, no preconditioner, convergence on the recurrence residual against
  max(rtol * ||b||, 1e-50), nonzero initial guess allowed.  PETSc 3.21.3 is not vendored either; these defaults are
  recalled from its documentation.
* ``CellVelocitySolver``  -- Compute_Rhs (ModVelSolver.F90:455-515), MyMatMult (:523-601) and Solve_RBC_Vel (:44-135)
  around two callables (single-layer and double-layer operator application), so that the same solve can be driven by
  the GPU library and by the CPU oracle.
"""
from __future__ import annotations

import numpy as np

from . import sphere

PHYS_TO_FOUR, FOUR_TO_PHYS = 1, 2


class ShTransform:
    """Gaussian-grid scalar SH analysis / synthesis truncated to degree < nlat0 (what Glob_Sph_Trans keeps)."""

    def __init__(self, nlat: int, nlon: int, nlat0: int):
        self.nlat, self.nlon, self.nlat0 = nlat, nlon, nlat0
        self.th, self.phi, w = sphere.gauss_grid(nlat, nlon)
        self.wg = w / (sphere.TWO_PI / nlon)                       # Gauss weights (sum = 2)
        self.pbar = sphere._pbar(nlat0, np.cos(self.th))           # (m, n, nlat), zero for n < m

    def anal(self, f: np.ndarray):
        """f (..., nlon, nlat) -> a, b (..., m, n) with 0 <= m <= n < nlat0."""
        F = np.fft.rfft(f, axis=-2)[..., :self.nlat0, :]           # sum_j f e^{-i m phi_j}
        # (1/pi) int f Pbar cos(m phi) dOmega = (2 / nlon) sum_i w_i Pbar(i) Re F_m(i)
        C = np.einsum("mni,...mi->...mn", self.pbar * self.wg, F) * (2.0 / self.nlon)
        return C.real.copy(), C.imag.copy()                        # b multiplies -sin: Im F = -sum f sin

    def synth(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        """a, b (..., m, n) -> f (..., nlon, nlat)."""
        # f = Re sum_m c_m (a + i b) Pbar e^{i m phi}, c_0 = 1/2; irfft weights: (1/nlon)(X_0 + 2 Re sum X_m e^{..})
        X = np.einsum("mni,...mn->...mi", self.pbar, a + 1j * b) * (self.nlon / 2.0)
        full = np.zeros(X.shape[:-2] + (self.nlon // 2 + 1, self.nlat), dtype=complex)
        full[..., :self.nlat0, :] = X
        full[..., 0, :] = full[..., 0, :].real
        return np.fft.irfft(full, n=self.nlon, axis=-2)


def _pack_index(nlat0: int):
    """order of Glob_Sph_Trans: a(m,n) for m = 0.., n = m..nlat0-1, then b(m,n) for m = 1.."""
    ia = [(m, n) for m in range(nlat0) for n in range(m, nlat0)]
    ib = [(m, n) for m in range(1, nlat0) for n in range(m, nlat0)]
    return np.array(ia).T, np.array(ib).T


class GlobSphTrans:
    """Glob_Sph_Trans for a suspension of identical meshes: v (3, ncell*nlon*nlat) <-> c (ncell*3*nlat0^2)."""

    def __init__(self, ncell: int, nlat: int, nlon: int, nlat0: int):
        self.ncell, self.nlat, self.nlon, self.nlat0 = ncell, nlat, nlon, nlat0
        self.sh = ShTransform(nlat, nlon, nlat0)
        self.ia, self.ib = _pack_index(nlat0)
        self.dof_cell = 3 * nlat0 * nlat0
        self.dof = ncell * self.dof_cell

    def phys_to_four(self, v: np.ndarray) -> np.ndarray:
        f = v.reshape(3, self.ncell, self.nlon, self.nlat)
        a, b = self.sh.anal(f)                                      # (3, ncell, m, n)
        ca = a[:, :, self.ia[0], self.ia[1]]                        # (3, ncell, na)
        cb = b[:, :, self.ib[0], self.ib[1]]
        c = np.concatenate([ca, cb], axis=2)                        # (3, ncell, nlat0^2)
        return np.ascontiguousarray(c.transpose(1, 2, 0)).reshape(-1)   # cell, coefficient, component

    def four_to_phys(self, c: np.ndarray) -> np.ndarray:
        cc = c.reshape(self.ncell, self.nlat0 * self.nlat0, 3).transpose(2, 0, 1)
        na = self.ia.shape[1]
        a = np.zeros((3, self.ncell, self.nlat0, self.nlat0))
        b = np.zeros_like(a)
        a[:, :, self.ia[0], self.ia[1]] = cc[:, :, :na]
        b[:, :, self.ib[0], self.ib[1]] = cc[:, :, na:]
        f = self.sh.synth(a, b)                                     # (3, ncell, nlon, nlat)
        return np.ascontiguousarray(f.reshape(3, -1))


def gmres(matvec, b, x0=None, rtol=1e-11, abstol=1e-50, restart=30, maxit=200):
    """KSPGMRES, PCNONE, classical Gram-Schmidt (no refinement).  Returns (x, niter, history) where history[k] is the
    residual norm PETSc would print after k iterations (history[0] = ||b - A x0||)."""
    b = np.asarray(b, dtype=float)
    x = np.zeros_like(b) if x0 is None else np.array(x0, dtype=float)
    bnorm = float(np.linalg.norm(b))
    ttol = max(rtol * bnorm, abstol)
    history = []
    it = 0
    while True:
        r = b - matvec(x) if (x0 is not None or it > 0) else b.copy()
        beta = float(np.linalg.norm(r))
        if it == 0:
            history.append(beta)
        if beta < ttol or it >= maxit:
            return x, it, history
        V = np.zeros((restart + 1, b.size))
        H = np.zeros((restart + 1, restart))
        cs, sn = np.zeros(restart), np.zeros(restart)
        gvec = np.zeros(restart + 1)
        gvec[0] = beta
        V[0] = r / beta
        k = 0
        while k < restart and it < maxit:
            w = matvec(V[k])
            h = V[:k + 1] @ w                      # classical Gram-Schmidt: all projections from the same w
            w = w - V[:k + 1].T @ h
            hn = float(np.linalg.norm(w))
            H[:k + 1, k] = h
            H[k + 1, k] = hn
            for i in range(k):                     # previous rotations
                t = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
                H[i + 1, k] = -sn[i] * H[i, k] + cs[i] * H[i + 1, k]
                H[i, k] = t
            d = np.hypot(H[k, k], H[k + 1, k])
            cs[k], sn[k] = H[k, k] / d, H[k + 1, k] / d
            H[k, k], H[k + 1, k] = d, 0.0
            gvec[k + 1] = -sn[k] * gvec[k]
            gvec[k] = cs[k] * gvec[k]
            it += 1
            k += 1
            res = abs(gvec[k])
            history.append(res)
            if hn > 0.0:
                V[k] = w / hn
            if res < ttol or hn == 0.0:
                break
        y = np.linalg.solve(np.triu(H[:k, :k]), gvec[:k])
        x = x + V[:k].T @ y
        if history[-1] < ttol or it >= maxit:
            return x, it, history
        x0 = x                                      # restart: true residual from the updated solution


class CellVelocitySolver:
    """Solve_RBC_Vel around the drop-in boundary.

    apply_sl(f_weighted) -> v (3, Np):   v = sum over cells of the single-layer integral, c1 = 1/(4 pi), c2 = 0
    apply_dl(g_weighted, g_raw) -> v:     the matvec operator, c1 = 0, c2 = -1/(4 pi)
    (weighted = density * detJ * w as SourceList_UpdateDensity stores it, raw = rbc%g on the mesh)
    """

    def __init__(self, sus, apply_sl, apply_dl):
        self.sus = sus
        self.apply_sl, self.apply_dl = apply_sl, apply_dl
        self.T = GlobSphTrans(sus.ncell, sus.nlat, sus.nlon, sus.nlat0)
        self.nmatvec = 0

    def compute_rhs(self, vbkg=(1.0, 0.0, 0.0)) -> np.ndarray:
        sus = self.sus
        v = self.apply_sl(sus.weighted(sus.f))
        A = np.repeat(sus.Acoef, sus.nlat * sus.nlon)
        v = v + 2.0 * np.asarray(vbkg, dtype=float)[:, None] / A[None, :]     # ModVelSolver.F90:497-500
        return self.T.phys_to_four(v)

    def matmult(self, u: np.ndarray) -> np.ndarray:
        g = self.T.four_to_phys(u)                                   # Glob_Sph_Trans(g, u, FOUR_TO_PHYS)
        v = self.apply_dl(self.sus.weighted(g), g)                   # off-diagonal term
        v = v + g                                                    # diagonal term, ModVelSolver.F90:587
        self.nmatvec += 1
        return self.T.phys_to_four(v)

    def solve(self, rhs=None, x0=None, rtol=1e-11, maxit=200, vbkg=(1.0, 0.0, 0.0)):
        """-> (sol coefficients, surface velocity (3, Np), niter, residual history)."""
        if rhs is None:
            rhs = self.compute_rhs(vbkg)
        if x0 is None:
            x0 = np.zeros_like(rhs)                                  # KSPSetInitialGuessNonzero with a zero vec_sol
        sol, niter, hist = gmres(self.matmult, rhs, x0=x0, rtol=rtol, maxit=maxit)
        return sol, self.T.four_to_phys(sol), niter, hist
