"""Host-side surface numerics that PRODUCE the operator's inputs (NumPy, init/step-time, not the hot path).

The Ewald operator consumes, per cell, mesh coordinates, normals, Jacobians and bicubic spline
coefficient arrays of x, a3, detJ, f*detJ and g*detJ.  In the reference those are produced by
SPHEREPACK + FFTW inside ``Rbc_BuildSurfaceSource`` (ModRbc.F90:716-807), ``Spline_Build_on_Sphere``
(ModSpline.F90:121-142) and ``FFT_Diff`` (ModFFT.F90:25-93), which stay on the Fortran side of the
drop-in boundary.  This module restates just enough of them to build synthetic suspensions and to
drive the tests/benchmarks without Fortran:

* Gauss colatitudes and weights (``gaqd`` as used in ModRbc.F90:93-95),
* scalar spherical-harmonic analysis on the Gauss grid, truncation to degree < nlat0 (``ShFilter``,
  ModSphpk.F90:423-432) and synthesis on the equally spaced colatitudes 0..pi (``ShSynthEqu``),
* the doubled-sphere periodic extension and spectral derivatives of ``Spline_Build_on_Sphere``,
* the analytic biconcave shape of ``RBC_MakeBiConcave`` (ModRbc.F90:368-398) and the surface
  geometry of ``RBC_ComputeGeometry`` (ModRbc.F90:404-509; a1 = dx/dtheta, a2 = dx/dphi,
  detJ = |a1 x a2| / sin(theta), a3 = unit normal).

Array conventions (shared with include/rbc3d.h and the oracle):
  mesh fields       (ncell, nvar, nlon, nlat)   -- ilat fastest, like Fortran x(nlat, nlon, nvar)
  spline arrays     (ncell, 4, nvar, nlon, 2*nlat) -- [u, u1, u2, u12], theta index fastest
"""
from __future__ import annotations

import numpy as np
from scipy.special import gammaln, lpmv

TWO_PI = 2.0 * np.pi


def gauss_grid(nlat: int, nlon: int):
    """th (ascending colatitudes), phi, w (Gauss weights * 2pi/nlon) -- ModRbc.F90:93-95."""
    xg, wg = np.polynomial.legendre.leggauss(nlat)
    th = np.arccos(xg[::-1]).copy()
    w = wg[::-1].copy() * (TWO_PI / nlon)
    phi = np.arange(nlon) * (TWO_PI / nlon)
    return th, phi, w


def _pbar(nmax: int, x: np.ndarray) -> np.ndarray:
    """Orthonormal associated Legendre functions Pbar[m, n, i] = N_nm P_n^m(x_i), 0 <= m <= n < nmax,
    normalised so that int_{-1}^{1} Pbar_n^m Pbar_n'^m dx = delta_nn'."""
    out = np.zeros((nmax, nmax, x.size))
    for m in range(nmax):
        for n in range(m, nmax):
            lognorm = 0.5 * (np.log(2 * n + 1.0) - np.log(2.0) + gammaln(n - m + 1) - gammaln(n + m + 1))
            out[m, n] = np.exp(lognorm) * lpmv(m, n, x)
    return out


class SphereProjector:
    """Gauss grid -> (filter to degree < nlat0) -> values on arbitrary colatitudes, same longitudes.

    One dense (n_out x nlat) matrix per zonal wavenumber m < nlat0; wavenumbers >= nlat0 are dropped
    (every degree n >= nlat0 is removed by ShFilter, and m <= n)."""

    def __init__(self, nlat: int, nlon: int, nlat0: int, th_out: np.ndarray):
        self.nlat, self.nlon, self.nlat0 = nlat, nlon, nlat0
        self.th, self.phi, self.w = gauss_grid(nlat, nlon)
        wg = self.w / (TWO_PI / nlon)
        pg = _pbar(nlat0, np.cos(self.th))          # (m, n, nlat)
        po = _pbar(nlat0, np.cos(np.asarray(th_out)))  # (m, n, nout)
        # T[m] = po[m].T @ (pg[m] * wg)
        self.T = np.einsum("mno,mni->moi", po, pg * wg[None, None, :])  # (m, nout, nlat)

    def __call__(self, f: np.ndarray) -> np.ndarray:
        """f (..., nlon, nlat) on the Gauss grid -> (..., nlon, nout)."""
        F = np.fft.rfft(f, axis=-2)                 # (..., nlon/2+1, nlat)
        nout = self.T.shape[1]
        G = np.zeros(F.shape[:-2] + (F.shape[-2], nout), dtype=complex)
        m = self.nlat0
        G[..., :m, :] = np.einsum("moi,...mi->...mo", self.T, F[..., :m, :])
        return np.fft.irfft(G, n=self.nlon, axis=-2)


def _fft_diff(u: np.ndarray, axis: int) -> np.ndarray:
    """Spectral derivative of a 2pi-periodic array along ``axis``, Nyquist mode zeroed -- FFT_Diff,
    ModFFT.F90:44-49."""
    n = u.shape[axis]
    U = np.fft.rfft(u, axis=axis)
    k = np.arange(n // 2 + 1, dtype=float)
    k[n // 2] = 0.0
    shape = [1] * u.ndim
    shape[axis] = k.size
    return np.fft.irfft(U * (1j * k).reshape(shape), n=n, axis=axis)


def spline_build_on_sphere(v: np.ndarray) -> np.ndarray:
    """v (..., nvar, nlon, nlat+1) on equally spaced colatitudes i*pi/nlat -> spline (..., 4, nvar, nlon, 2*nlat).

    Spline_Build_on_Sphere, ModSpline.F90:121-142: u(i,j) = v(i,j), u(nlat+i, j) = v(nlat-i, j+nlon/2)."""
    nlat = v.shape[-1] - 1
    nlon = v.shape[-2]
    u = np.empty(v.shape[:-1] + (2 * nlat,))
    u[..., :nlat] = v[..., :nlat]
    u[..., nlat:] = np.roll(v, -(nlon // 2), axis=-2)[..., nlat:0:-1]
    u1 = _fft_diff(u, -1)
    u2 = _fft_diff(u, -2)
    u12 = _fft_diff(u1, -2)
    return np.stack([u, u1, u2, u12], axis=-4)


class SurfaceSplines:
    """Rbc_BuildSurfaceSource (ModRbc.F90:716-807): mesh field on the Gauss grid -> spline arrays."""

    def __init__(self, nlat0: int, dealias: int = 3):
        self.nlat0 = nlat0
        self.nlat = dealias * nlat0
        self.nlon = 2 * self.nlat
        th_equ = np.arange(self.nlat + 1) * (np.pi / self.nlat)
        self.proj = SphereProjector(self.nlat, self.nlon, nlat0, th_equ)

    def build(self, f: np.ndarray) -> np.ndarray:
        """f (ncell, nvar, nlon, nlat) -> (ncell, 4, nvar, nlon, 2*nlat), C-contiguous."""
        return np.ascontiguousarray(spline_build_on_sphere(self.proj(f)))


def biconcave_unit(th: np.ndarray, phi: np.ndarray, rad: float = 1.0):
    """RBC_MakeBiConcave (ModRbc.F90:368-398) and its analytic tangent vectors.

    Returns x, a1 = dx/dtheta, a2 = dx/dphi, each (3, nlon, nlat), in the cell's own frame."""
    alph = 1.3858189
    T, Ph = np.meshgrid(th, phi, indexing="xy")     # (nlon, nlat)
    s, c = np.sin(T), np.cos(T)
    poly = 0.207 + 2.003 * s ** 2 - 1.123 * s ** 4
    dpoly = (2 * 2.003 * s - 4 * 1.123 * s ** 3) * c
    z = rad * 0.5 * alph * poly * c
    dz = rad * 0.5 * alph * (dpoly * c - poly * s)
    r = rad * alph * s
    dr = rad * alph * c
    x = np.stack([r * np.cos(Ph), r * np.sin(Ph), z])
    a1 = np.stack([dr * np.cos(Ph), dr * np.sin(Ph), dz])
    a2 = np.stack([-r * np.sin(Ph), r * np.cos(Ph), np.zeros_like(z)])
    return x, a1, a2


def _dpbar_dth(nmax: int, th: np.ndarray) -> np.ndarray:
    """d Pbar_n^m(cos th) / d th for 0 <= m <= n < nmax at colatitudes th (no pole among them):
    sin(th) dP_n^m/dth = n cos(th) P_n^m - (n + m) P_{n-1}^m, carried over to the orthonormal functions."""
    x, s = np.cos(th), np.sin(th)
    out = np.zeros((nmax, nmax, th.size))
    for m in range(nmax):
        for n in range(m, nmax):
            lognorm = 0.5 * (np.log(2 * n + 1.0) - np.log(2.0) + gammaln(n - m + 1) - gammaln(n + m + 1))
            pn = lpmv(m, n, x)
            pn1 = lpmv(m, n - 1, x) if n - 1 >= m else np.zeros_like(x)
            out[m, n] = np.exp(lognorm) * (n * x * pn - (n + m) * pn1) / s
    return out


class SphereGradient:
    """ShAnalGau + ShGradGau (ModSphpk.F90:74-105, 377-417; SPHEREPACK shags + gradgs, then the sin(th) factor put
    back): scalar fields on the nlat x nlon Gauss grid -> (d/d theta, d/d phi), all degrees n < nlat kept."""

    def __init__(self, nlat: int, nlon: int):
        self.nlat, self.nlon = nlat, nlon
        self.th, self.phi, w = gauss_grid(nlat, nlon)
        wg = w / (TWO_PI / nlon)
        mmax = min(nlat, nlon // 2 + 1)
        pb = _pbar(nlat, np.cos(self.th))[:mmax]                   # (m, n, nlat)
        dpb = _dpbar_dth(nlat, self.th)[:mmax]
        self.mmax = mmax
        self.D = np.einsum("mno,mni->moi", dpb, pb * wg[None, None, :])   # values -> d/dth, per zonal wavenumber

    def __call__(self, f: np.ndarray):
        """f (..., nlon, nlat) -> f_theta, f_phi (..., nlon, nlat)."""
        F = np.fft.rfft(f, axis=-2)
        G = np.zeros_like(F)
        G[..., :self.mmax, :] = np.einsum("moi,...mi->...mo", self.D, F[..., :self.mmax, :])
        m = np.arange(F.shape[-2])
        Fp = 1j * m[:, None] * F
        if self.nlon % 2 == 0:
            Fp[..., -1, :] = 0.0
        return np.fft.irfft(G, n=self.nlon, axis=-2), np.fft.irfft(Fp, n=self.nlon, axis=-2)


def surface_geometry(a1: np.ndarray, a2: np.ndarray, th: np.ndarray):
    """a3 (unit normal) and detJ = |a1 x a2| / sin(theta) -- ModRbc.F90:431-455.
    a1, a2 (..., 3, nlon, nlat)."""
    cr = np.cross(a1, a2, axis=-3)
    nrm = np.sqrt((cr ** 2).sum(axis=-3))
    a3 = cr / nrm[..., None, :, :]
    detj = nrm / np.sin(th)
    return a3, detj


def rotate_matrix(c1) -> np.ndarray:
    """RotateMatrix (ModBasicMath.F90:97-126): the rotation that takes the z axis to the unit vector c1 about the axis
    a = z x c1; mat = a a^T + b1 b^T + c1 z^T with b = z x a, b1 = c1 x a; the identity when c1 is (anti)parallel to z."""
    c1 = np.asarray(c1, dtype=float)
    c = np.array([0.0, 0.0, 1.0])
    a = np.array([-c1[1], c1[0], 0.0])
    if (a * a).sum() < 1e-10:
        return np.eye(3)
    a = a / np.sqrt((a * a).sum())
    b = np.array([-a[1], a[0], 0.0])
    b1 = np.cross(c1, a)
    return np.outer(a, a) + np.outer(b1, b) + np.outer(c1, c)


def rotation_matrices(rng: np.random.Generator, n: int) -> np.ndarray:
    """n uniformly random rotations (n, 3, 3) from unit quaternions."""
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    a, b, c, d = q.T
    return np.stack([
        np.stack([a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)], -1),
        np.stack([2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)], -1),
        np.stack([2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d], -1),
    ], 1)


def random_bandlimited_field(rng: np.random.Generator, ncell: int, nvar: int, nlat0: int, th: np.ndarray,
                             nlon: int) -> np.ndarray:
    """Band-limited random field (SH degree < nlat0, coefficients ~ U(-1,1)) on the Gauss grid,
    shape (ncell, nvar, nlon, nlat)."""
    pg = _pbar(nlat0, np.cos(th))                   # (m, n, nlat)
    ca = rng.uniform(-1, 1, size=(ncell, nvar, nlat0, nlat0))
    cb = rng.uniform(-1, 1, size=(ncell, nvar, nlat0, nlat0))
    F = np.zeros((ncell, nvar, nlon // 2 + 1, th.size), dtype=complex)
    F[:, :, :nlat0, :] = np.einsum("cvmn,mni->cvmi", ca + 1j * cb, pg)
    F[:, :, 0, :] = F[:, :, 0, :].real
    return np.fft.irfft(F, n=nlon, axis=-2) * nlon / (2 * np.sqrt(np.pi) * nlat0)
