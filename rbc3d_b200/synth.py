"""Synthetic periodic suspensions of biconcave cells (BASELINE.json configs[2]).

nrbc = n^3 cells on a cubic lattice in a cubic box of side L = 57.2588 * n / 16 (so that the 4096-cell
case gets the 256^3 PME mesh and the 47^3 cell list at alpha = 0.44, eps = 1e-3, P = 8), every cell the
analytic biconcave shape of ``RBC_MakeBiConcave`` (ModRbc.F90:368-398) with equivalent radius 1, randomly
rotated and jittered by +-0.2; viscosity ratio 5 (Acoef = 6, Bcoef = -4, ModConf.F90:333-340) so that the
double-layer matvec of ModVelSolver.F90:523-601 exists; band-limited random densities.
RNG: numpy PCG64(seed), default seed 161269 (the reference's ``ranseed``, examples/minicase/minit.F90:25).

Only NumPy; nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import sphere

# SURVEY.md quotes 57.2589, but 2*ceil(sqrt(-ln eps/(pi alpha))*57.2589) = 258 (the product is 128.00007);
# 57.2588 is the nearest 4-decimal side that yields the 256^3 mesh BASELINE.json names (and still Nc = 47).
L_4096 = 57.2588


@dataclass
class Suspension:
    Lb: np.ndarray
    nlat0: int
    nlat: int
    nlon: int
    ncell: int
    th: np.ndarray
    phi: np.ndarray
    w: np.ndarray
    # flat point lists, SoA (3, Np), p = cell*nlat*nlon + ilon*nlat + ilat
    x: np.ndarray
    a3: np.ndarray
    detj: np.ndarray          # (Np,)
    f: np.ndarray             # raw single-layer density rbc%f
    g: np.ndarray             # raw double-layer density rbc%g
    Acoef: np.ndarray         # (ncell,)
    Bcoef: np.ndarray         # (ncell,)
    area: np.ndarray          # (ncell,)
    meshSize: np.ndarray      # (ncell,)
    spx: np.ndarray           # (ncell, 4, 3, nlon, 2 nlat)
    spa3: np.ndarray
    spdetj: np.ndarray        # (ncell, 4, 1, nlon, 2 nlat)
    spF: np.ndarray | None
    spG: np.ndarray | None
    centers: np.ndarray = field(default=None)

    @property
    def npoint(self) -> int:
        return self.ncell * self.nlat * self.nlon

    def dS(self) -> np.ndarray:
        """detJ * w per point: the quadrature weight of SourceList_UpdateDensity (ModSourceList.F90:181)."""
        return self.detj * np.tile(self.w, self.ncell * self.nlon)

    def weighted(self, dens: np.ndarray) -> np.ndarray:
        """slist%f / slist%g as the source list holds them (ModSourceList.F90:183-184)."""
        return dens * self.dS()[None, :]


def _flat(field_: np.ndarray) -> np.ndarray:
    """(ncell, nvar, nlon, nlat) -> (nvar, Np)"""
    nc, nv = field_.shape[:2]
    return np.ascontiguousarray(field_.transpose(1, 0, 2, 3).reshape(nv, -1))


def build_splines(sus: Suspension, builder: "sphere.SurfaceSplines", which=("x", "a3", "detj", "F", "G"),
                  chunk: int = 128) -> None:
    """(Re)build spline coefficient arrays from the mesh fields, as Rbc_BuildSurfaceSource does."""
    nc, nlat, nlon = sus.ncell, sus.nlat, sus.nlon

    def mesh(a, nv):
        return a.reshape(nv, nc, nlon, nlat).transpose(1, 0, 2, 3)

    dj = sus.detj.reshape(nc, 1, nlon, nlat)
    jobs = []
    if "x" in which:
        jobs.append(("spx", mesh(sus.x, 3), 3))
    if "a3" in which:
        jobs.append(("spa3", mesh(sus.a3, 3), 3))
    if "detj" in which:
        jobs.append(("spdetj", dj, 1))
    if "F" in which and sus.f is not None:
        jobs.append(("spF", mesh(sus.f, 3) * dj, 3))
    if "G" in which and sus.g is not None:
        jobs.append(("spG", mesh(sus.g, 3) * dj, 3))
    for name, src, nv in jobs:
        out = np.empty((nc, 4, nv, nlon, 2 * nlat))
        for c0 in range(0, nc, chunk):
            out[c0:c0 + chunk] = builder.build(np.ascontiguousarray(src[c0:c0 + chunk]))
        setattr(sus, name, out)


def make_suspension(n_side: int = 4, nlat0: int = 12, dealias: int = 3, seed: int = 161269,
                    visc_ratio: float = 5.0, jitter: float = 0.2, L: float | None = None,
                    spacing_scale: float = 1.0, with_f: bool = True, with_g: bool = True,
                    centers: np.ndarray | None = None) -> Suspension:
    """Build an n_side^3-cell suspension.  ``spacing_scale`` < 1 packs the lattice tighter (used by the
    tests to force near-singular cell-cell pairs); ``centers`` overrides the lattice (cells, 3)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    builder = sphere.SurfaceSplines(nlat0, dealias)
    nlat, nlon = builder.nlat, builder.nlon
    th, phi, w = sphere.gauss_grid(nlat, nlon)
    if L is None:
        L = L_4096 * n_side / 16.0 * spacing_scale
    Lb = np.array([L, L, L], dtype=float)
    if centers is None:
        idx = np.stack(np.meshgrid(*[np.arange(n_side)] * 3, indexing="ij"), -1).reshape(-1, 3)
        centers = (idx + 0.5) * (L / n_side) + rng.uniform(-jitter, jitter, size=idx.shape)
    centers = np.asarray(centers, dtype=float)
    ncell = centers.shape[0]
    R = sphere.rotation_matrices(rng, ncell)

    xu, a1u, a2u = sphere.biconcave_unit(th, phi, 1.0)          # (3, nlon, nlat)
    x = np.einsum("cij,jlk->cilk", R, xu) + centers[:, :, None, None]
    a1 = np.einsum("cij,jlk->cilk", R, a1u)
    a2 = np.einsum("cij,jlk->cilk", R, a2u)
    a3, detj = sphere.surface_geometry(a1, a2, th)
    ds = detj * w                                             # (ncell, nlon, nlat)
    area = ds.sum(axis=(1, 2))
    meshSize = np.sqrt(area / nlat ** 2)                       # ModRbc.F90:504

    f = g = None
    if with_f:
        f = _flat(sphere.random_bandlimited_field(rng, ncell, 3, nlat0, th, nlon))
    if with_g:
        g = _flat(sphere.random_bandlimited_field(rng, ncell, 3, nlat0, th, nlon))
    A = np.full(ncell, 1.0 + visc_ratio)
    B = np.full(ncell, 1.0 - visc_ratio)
    sus = Suspension(Lb=Lb, nlat0=nlat0, nlat=nlat, nlon=nlon, ncell=ncell, th=th, phi=phi, w=w,
                     x=_flat(x), a3=_flat(a3), detj=np.ascontiguousarray(detj.reshape(-1)), f=f, g=g,
                     Acoef=A, Bcoef=B, area=area, meshSize=meshSize, spx=None, spa3=None, spdetj=None,
                     spF=None, spG=None, centers=centers)
    build_splines(sus, builder)
    sus._builder = builder
    return sus
