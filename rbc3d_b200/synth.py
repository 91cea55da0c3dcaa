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
                    centers: np.ndarray | None = None, rotate: bool = True) -> Suspension:
    """Build an n_side^3-cell suspension.  ``spacing_scale`` < 1 packs the lattice tighter (used by the
    tests to force near-singular cell-cell pairs); ``centers`` overrides the lattice (cells, 3); ``rotate=False``
    keeps every cell in the orientation of RBC_MakeBiConcave (axis along z), as the example init programs do."""
    rng = np.random.Generator(np.random.PCG64(seed))
    builder = sphere.SurfaceSplines(nlat0, dealias)
    nlat, nlon = builder.nlat, builder.nlon
    th, phi, w = sphere.gauss_grid(nlat, nlon)
    if L is None:
        L = L_4096 * n_side / 16.0 * spacing_scale
    Lb = np.asarray(L, dtype=float) * np.ones(3)      # scalar (cubic box) or the three box lengths
    if centers is None:
        idx = np.stack(np.meshgrid(*[np.arange(n_side)] * 3, indexing="ij"), -1).reshape(-1, 3)
        centers = (idx + 0.5) * (Lb / n_side) + rng.uniform(-jitter, jitter, size=idx.shape)
    centers = np.asarray(centers, dtype=float)
    ncell = centers.shape[0]
    R = sphere.rotation_matrices(rng, ncell)
    if not rotate:
        R = np.broadcast_to(np.eye(3), (ncell, 3, 3)).copy()

    xu, a1u, a2u = sphere.biconcave_unit(th, phi, 1.0)          # (3, nlon, nlat)
    x = np.einsum("cij,jlk->cilk", R, xu) + centers[:, :, None, None]
    a1 = np.einsum("cij,jlk->cilk", R, a1u)
    a2 = np.einsum("cij,jlk->cilk", R, a2u)
    return _assemble(rng, builder, Lb, nlat0, x, a1, a2, np.full(ncell, float(visc_ratio)), centers, with_f, with_g)


def _assemble(rng, builder, Lb, nlat0, x, a1, a2, visc_ratio, centers, with_f, with_g) -> Suspension:
    """x, a1, a2 (ncell, 3, nlon, nlat) -> Suspension with normals, Jacobians, densities, coefficients and splines."""
    nlat, nlon = builder.nlat, builder.nlon
    th, phi, w = sphere.gauss_grid(nlat, nlon)
    ncell = x.shape[0]
    a3, detj = sphere.surface_geometry(a1, a2, th)
    ds = detj * w                                             # (ncell, nlon, nlat)
    area = ds.sum(axis=(1, 2))
    meshSize = np.sqrt(area / nlat ** 2)                       # ModRbc.F90:504

    f = g = None
    if with_f:
        f = _flat(sphere.random_bandlimited_field(rng, ncell, 3, nlat0, th, nlon))
    if with_g:
        g = _flat(sphere.random_bandlimited_field(rng, ncell, 3, nlat0, th, nlon))
    A = 1.0 + np.asarray(visc_ratio, dtype=float)              # initCOEFs: A = 1 + lambda, B = 1 - lambda
    B = 1.0 - np.asarray(visc_ratio, dtype=float)
    sus = Suspension(Lb=Lb, nlat0=nlat0, nlat=nlat, nlon=nlon, ncell=ncell, th=th, phi=phi, w=w,
                     x=_flat(x), a3=_flat(a3), detj=np.ascontiguousarray(detj.reshape(-1)), f=f, g=g,
                     Acoef=A, Bcoef=B, area=area, meshSize=meshSize, spx=None, spa3=None, spdetj=None,
                     spF=None, spG=None, centers=centers)
    build_splines(sus, builder)
    sus._builder = builder
    return sus


def suspension_from_shapes(x: np.ndarray, Lb, nlat0: int = 12, dealias: int = 3, visc_ratio=1.0, seed: int = 161269,
                           with_f: bool = True, with_g: bool = True) -> Suspension:
    """Cells given by their mesh coordinates x (ncell, 3, nlon, nlat) -- imported shapes (ImportReadRBC) next to analytic
    ones: tangent vectors by spectral differentiation as RBC_ComputeGeometry does (ModRbc.F90:419-456: ShAnalGau +
    ShGradGau over all degrees of the nlat x nlon grid).  ``visc_ratio``: scalar or one value per cell."""
    rng = np.random.Generator(np.random.PCG64(seed))
    builder = sphere.SurfaceSplines(nlat0, dealias)
    x = np.ascontiguousarray(x, dtype=float)
    ncell = x.shape[0]
    assert x.shape == (ncell, 3, builder.nlon, builder.nlat)
    a1, a2 = sphere.SphereGradient(builder.nlat, builder.nlon)(x)
    lam = np.broadcast_to(np.asarray(visc_ratio, dtype=float), (ncell,)).copy()
    Lb = np.asarray(Lb, dtype=float) * np.ones(3)
    return _assemble(rng, builder, Lb, nlat0, x, a1, a2, lam, x.mean(axis=(2, 3)), with_f, with_g)


def subset(sus: Suspension, cells) -> Suspension:
    """The suspension restricted to the given cells (in the given order) -- what a rank holds after the reference's
    source filter (Cell_Has_Source, ModConf.F90:467-493; rbc3d_b200/partition.py)."""
    cells = np.asarray(cells, dtype=np.int64)
    npc = sus.nlat * sus.nlon
    pts = (cells[:, None] * npc + np.arange(npc)[None, :]).reshape(-1)
    pick = lambda a: None if a is None else np.ascontiguousarray(a[..., pts])      # noqa: E731
    cpick = lambda a: None if a is None else np.ascontiguousarray(a[cells])        # noqa: E731
    out = Suspension(Lb=sus.Lb, nlat0=sus.nlat0, nlat=sus.nlat, nlon=sus.nlon, ncell=len(cells), th=sus.th, phi=sus.phi,
                     w=sus.w, x=pick(sus.x), a3=pick(sus.a3), detj=pick(sus.detj), f=pick(sus.f), g=pick(sus.g),
                     Acoef=cpick(sus.Acoef), Bcoef=cpick(sus.Bcoef), area=cpick(sus.area), meshSize=cpick(sus.meshSize),
                     spx=cpick(sus.spx), spa3=cpick(sus.spa3), spdetj=cpick(sus.spdetj), spF=cpick(sus.spF),
                     spG=cpick(sus.spG), centers=cpick(sus.centers))
    out._builder = getattr(sus, "_builder", None)
    return out


# ---------------------------------------------------------------------------------------------------------
# Walls: triangulated tubes like the vessel of examples/minicase and examples/case (a cylinder along z whose end
# rings sit on z = 0 and z = Lb3 as duplicated vertices, the way the reference's Exodus meshes close the period).
@dataclass
class Walls:
    nvert: np.ndarray         # (nwall,) int32
    nele: np.ndarray          # (nwall,) int32
    x: np.ndarray             # SoA (3, NV): vertices of all walls back to back (= tlist_wall%x)
    e2v: np.ndarray           # SoA (3, NE) int32: wall%e2v, 1-based vertex numbers local to the wall
    area: np.ndarray          # (NE,)  Wall_ComputeGeometry, ModWall.F90:118-145
    epsDist: np.ndarray       # (NE,)  sqrt(area)
    f: np.ndarray             # SoA (3, NV) tractions wall%f

    @property
    def nwall(self) -> int:
        return len(self.nvert)

    @property
    def NV(self) -> int:
        return int(self.nvert.sum())

    @property
    def NE(self) -> int:
        return int(self.nele.sum())

    def voff(self):
        return np.concatenate([[0], np.cumsum(self.nvert)]).astype(np.int64)

    def eoff(self):
        return np.concatenate([[0], np.cumsum(self.nele)]).astype(np.int64)

    def e2v_global(self) -> np.ndarray:
        """(3, NE) 0-based indices into the concatenated vertex list."""
        out = self.e2v.astype(np.int64) - 1
        vo, eo = self.voff(), self.eoff()
        for w in range(self.nwall):
            out[:, eo[w]:eo[w + 1]] += vo[w]
        return out


def wall_geometry(x: np.ndarray, e2v_glb: np.ndarray):
    """area and epsDist of every element with the arithmetic of Wall_ComputeGeometry (ModWall.F90:118-145)."""
    xe = x[:, e2v_glb]                      # (3 comps, 3 corners, NE)
    x12, x13 = xe[:, 1] - xe[:, 0], xe[:, 2] - xe[:, 0]
    a3 = np.stack([x12[1] * x13[2] - x12[2] * x13[1], x12[2] * x13[0] - x12[0] * x13[2],
                   x12[0] * x13[1] - x12[1] * x13[0]])
    a3n = np.sqrt(a3[0] * a3[0] + a3[1] * a3[1] + a3[2] * a3[2])
    area = 0.5 * a3n
    return area, np.sqrt(area)


def tube_mesh(Lz: float, radius: float, ntheta: int, nz: int, center=(0.0, 0.0), z0: float = 0.0, wobble: float = 0.0,
              rng=None):
    """Vertices (3, (nz+1) ntheta) and 1-based e2v (3, 2 nz ntheta) of a cylinder along z, rings at z0 + k Lz/nz,
    k = 0..nz (first and last ring are periodic duplicates)."""
    k, j = np.meshgrid(np.arange(nz + 1), np.arange(ntheta), indexing="ij")
    ang = 2 * np.pi * (j + 0.5 * (k % 2)) / ntheta
    r = radius * np.ones_like(ang)
    if wobble and rng is not None:
        dr = wobble * rng.uniform(-1, 1, size=(nz, ntheta))
        r[:nz] += dr
        r[nz] = r[0]                        # the duplicated ring must coincide with ring 0 modulo the period
    x = np.stack([center[0] + r * np.cos(ang), center[1] + r * np.sin(ang), z0 + Lz * k / nz]).reshape(3, -1)
    # ring nz duplicates ring 0 only when nz is even (the half-cell stagger); force it
    if nz % 2:
        raise ValueError("nz must be even so that the staggered end rings coincide")
    vid = lambda kk, jj: kk * ntheta + (jj % ntheta)  # noqa: E731
    tri = []
    for kk in range(nz):
        for jj in range(ntheta):
            a, b, c, d = vid(kk, jj), vid(kk, jj + 1), vid(kk + 1, jj), vid(kk + 1, jj + 1)
            if kk % 2 == 0:
                tri += [(a, b, c), (b, d, c)]
            else:
                tri += [(a, b, d), (a, d, c)]
    e2v = np.array(tri, dtype=np.int32).T + 1
    return x, np.ascontiguousarray(e2v)


def make_walls(Lb, specs, seed: int = 161269, wobble: float = 0.0) -> Walls:
    """specs: list of dicts(radius, ntheta, nz[, center]) -- one tube wall each, spanning the z period of the box.
    Tractions: smooth random field (low-order trigonometric polynomial in angle and z) per wall."""
    rng = np.random.default_rng(np.random.PCG64(seed + 1))
    xs, es, nv, ne, fs = [], [], [], [], []
    for s in specs:
        x, e2v = tube_mesh(float(Lb[2]), s["radius"], s["ntheta"], s["nz"], center=s.get("center", (0.5 * Lb[0], 0.5 * Lb[1])),
                           wobble=wobble, rng=rng)
        xs.append(x)
        es.append(e2v)
        nv.append(x.shape[1])
        ne.append(e2v.shape[1])
        cx, cy = s.get("center", (0.5 * Lb[0], 0.5 * Lb[1]))
        ang = np.arctan2(x[1] - cy, x[0] - cx)
        zz = 2 * np.pi * x[2] / Lb[2]
        f = np.zeros_like(x)
        for d in range(3):
            for m in range(3):
                a, b, c, e = rng.uniform(-1, 1, 4)
                f[d] += a * np.cos(m * ang) * np.cos(m * zz) + b * np.sin(m * ang) + c * np.sin(m * zz) + e * np.cos(m * ang + zz)
        fs.append(f)
    x = np.ascontiguousarray(np.concatenate(xs, axis=1))
    e2v = np.ascontiguousarray(np.concatenate(es, axis=1))
    W = Walls(np.array(nv, np.int32), np.array(ne, np.int32), x, e2v, None, None, np.ascontiguousarray(np.concatenate(fs, axis=1)))
    W.area, W.epsDist = wall_geometry(x, W.e2v_global())
    return W
