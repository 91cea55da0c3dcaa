"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE -- see oracle/rbc3d_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
PARITY UNPINNED: the reference has no golden vectors and cannot be built here (see the header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "librbc3d_oracle.so")
NTAB = 8192

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("rbc3d_oracle.c", "rbc3d_oracle_walls.c", "rbc3d_oracle.h")]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


class Params(C.Structure):
    _fields_ = [("Lb", C.c_double * 3), ("iLb", C.c_double * 3), ("alpha", C.c_double), ("eps", C.c_double),
                ("rc", C.c_double), ("P", C.c_int), ("Nb", C.c_int * 3), ("Nc", C.c_int * 3),
                ("iLbNc", C.c_double * 3), ("sl_c1", C.c_double * (NTAB + 1)), ("sl_c2", C.c_double * (NTAB + 1)),
                ("dl_c1", C.c_double * (NTAB + 1)), ("mask_tab", C.c_double * (NTAB + 1)), ("r_eps", C.c_double)]


class Cells(C.Structure):
    _fields_ = [("ncell", C.c_int), ("nlat", C.c_int), ("nlon", C.c_int),
                ("th", c_dp), ("phi", c_dp), ("w", c_dp),
                ("x", c_dp), ("a3", c_dp), ("f", c_dp), ("g", c_dp), ("detj", c_dp),
                ("Acoef", c_dp), ("Bcoef", c_dp), ("area", c_dp), ("meshSize", c_dp),
                ("spx", c_dp), ("spa3", c_dp), ("spdetj", c_dp), ("spF", c_dp), ("spG", c_dp),
                ("patch_radius", C.c_double), ("nrad", C.c_int), ("nazm", C.c_int),
                ("thG", c_dp), ("phiG", c_dp), ("patch_w", c_dp)]


class WallsStruct(C.Structure):
    _fields_ = [("nwall", C.c_int), ("nvert", c_ip), ("nele", c_ip), ("id0", C.c_int), ("x", c_dp), ("f", c_dp),
                ("e2v", c_ip), ("area", c_dp), ("epsDist", c_dp)]


class Targets(C.Structure):
    _fields_ = [("n", C.c_int), ("x", c_dp), ("Acoef", c_dp), ("indx", c_ip), ("active", c_ip)]


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_ip)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.orc_pme_init.restype = C.c_void_p
        L.orc_pme_vv.restype = c_dp
        L.orc_pme_bb.restype = c_dp
        L.orc_mask_func.restype = C.c_double
        L.orc_mask_func_exact.restype = C.c_double
        L.orc_dist_on_sphere.restype = C.c_double
        L.orc_num_threads.restype = C.c_int
        L.orc_min_dist_to_tri.restype = C.c_double
        L.orc_prepare_sing_int_on_wall.restype = C.c_void_p
        L.orc_wallmat_nblk.restype = C.c_int
        _lib = L
    return _lib


class Oracle:
    """Stateful wrapper: parameters + (optionally) cells + PME state, like the reference's module globals."""

    FLAG_NO_SING, FLAG_NO_NEARSING, FLAG_NO_LINEAR, FLAG_NO_PAIRS = 1, 2, 4, 8

    def __init__(self, Lb, alpha=0.44, eps=1e-3, P=8, nranks=1, rc=None, Nb=None):
        L = lib()
        Lb3 = (C.c_double * 3)(*[float(v) for v in Lb])
        rc_c = C.c_double()
        Nb_c = (C.c_int * 3)()
        L.orc_set_ewald_prms(Lb3, C.c_double(alpha), C.c_double(eps), C.c_int(P), C.c_int(nranks), C.byref(rc_c), Nb_c)
        self.rc = float(rc_c.value) if rc is None else float(rc)
        self.Nb = [int(v) for v in Nb_c] if Nb is None else [int(v) for v in Nb]
        self.prm = Params()
        L.orc_params_init(C.byref(self.prm), Lb3, C.c_double(alpha), C.c_double(eps), C.c_int(P),
                          C.c_double(self.rc), (C.c_int * 3)(*self.Nb))
        self.Lb = np.array(Lb, dtype=float)
        self.alpha, self.eps, self.P = alpha, eps, P
        self.Nc = [int(v) for v in self.prm.Nc]
        self._pme = None
        self._keep = []
        self.cells = None

    # ---- scalar helpers -------------------------------------------------
    def ewald_sl(self, r):
        a, b = C.c_double(), C.c_double()
        lib().orc_ewald_coeff_sl(C.byref(self.prm), C.c_double(r), C.byref(a), C.byref(b))
        return a.value, b.value

    def ewald_dl(self, r):
        a = C.c_double()
        lib().orc_ewald_coeff_dl(C.byref(self.prm), C.c_double(r), C.byref(a))
        return a.value

    @staticmethod
    def ewald_sl_exact(r, alpha):
        a, b = C.c_double(), C.c_double()
        lib().orc_ewald_coeff_sl_exact(C.c_double(r), C.c_double(alpha), C.byref(a), C.byref(b))
        return a.value, b.value

    @staticmethod
    def ewald_dl_exact(r, alpha):
        a = C.c_double()
        lib().orc_ewald_coeff_dl_exact(C.c_double(r), C.c_double(alpha), C.byref(a))
        return a.value

    def mask(self, x):
        return lib().orc_mask_func(C.byref(self.prm), C.c_double(x))

    @staticmethod
    def bspline(xc, P):
        w = np.zeros(P)
        imin = C.c_int()
        lib().orc_bspline_func(C.c_double(xc), C.c_int(P), C.byref(imin), _dp(w))
        return imin.value, w

    @staticmethod
    def gauleg(x1, x2, n):
        x, w = np.zeros(n), np.zeros(n)
        lib().orc_gauleg(C.c_double(x1), C.c_double(x2), C.c_int(n), _dp(x), _dp(w))
        return x, w

    @staticmethod
    def gauleg_sinh(xmin, xmax, a, b, n):
        x, w = np.zeros(n), np.zeros(n)
        lib().orc_gauleg_sinh(C.c_double(xmin), C.c_double(xmax), C.c_double(a), C.c_double(b), C.c_int(n), _dp(x), _dp(w))
        return x, w

    @staticmethod
    def quadfit_2d(xy, f):
        xy = _f64(xy)
        f = _f64(f)
        c = np.zeros(6)
        info = lib().orc_quadfit_2d(C.c_int(len(f)), _dp(xy), _dp(f), _dp(c))
        return info, c

    def polar_patch(self, nlat, nlon, th, phi):
        radius, nrad, nazm = C.c_double(), C.c_int(), C.c_int()
        th, phi = _f64(th), _f64(phi)
        lib().orc_rbc_polar_patch_create(C.byref(self.prm), nlat, nlon, _dp(th), _dp(phi), C.byref(radius),
                                         C.byref(nrad), C.byref(nazm), None, None, None)
        thG = np.zeros((nlon, nlat, nazm.value, nrad.value))
        phiG = np.zeros_like(thG)
        w = np.zeros(nrad.value)
        lib().orc_rbc_polar_patch_create(C.byref(self.prm), nlat, nlon, _dp(th), _dp(phi), C.byref(radius),
                                         C.byref(nrad), C.byref(nazm), _dp(thG), _dp(phiG), _dp(w))
        return radius.value, nrad.value, nazm.value, thG, phiG, w

    @staticmethod
    def spline_interp(sp, x, y):
        sp = _f64(sp)
        _, nvar, n, m = sp.shape
        f = np.zeros(nvar)
        lib().orc_spline_interp(_dp(sp), C.c_int(m), C.c_int(n), C.c_int(nvar), C.c_double(x), C.c_double(y), _dp(f))
        return f

    @staticmethod
    def find_projection(sp, xtar, th0, phi0):
        sp = _f64(sp)
        _, nvar, n, m = sp.shape
        xt = _f64(xtar)
        t, p = C.c_double(th0), C.c_double(phi0)
        x0 = np.zeros(3)
        lib().orc_spline_find_projection(_dp(sp), C.c_int(m), C.c_int(n), _dp(xt), C.byref(t), C.byref(p), _dp(x0))
        return t.value, p.value, x0

    # ---- cell list --------------------------------------------------------
    def cell_ids(self, x):
        x = _f64(x)
        n = x.shape[1]
        cid = np.zeros(n, dtype=np.int32)
        lib().orc_cell_ids(C.byref(self.prm), C.c_int(n), _dp(x), _ip(cid))
        return cid

    def neighbor_signature(self, xs, xt):
        xs, xt = _f64(xs), _f64(xt)
        ns, nt = xs.shape[1], xt.shape[1]
        cnt = np.zeros(nt, dtype=np.int32)
        sig = np.zeros(nt, dtype=np.uint64)
        lib().orc_neighbor_signature(C.byref(self.prm), C.c_int(ns), _dp(xs), C.c_int(nt), _dp(xt), _ip(cnt),
                                     sig.ctypes.data_as(C.POINTER(C.c_ulonglong)))
        return cnt, sig

    # ---- cells ------------------------------------------------------------
    def set_cells(self, sus):
        """sus: rbc3d_b200.synth.Suspension (or any object with the same array attributes)."""
        radius, nrad, nazm, thG, phiG, pw = self.polar_patch(sus.nlat, sus.nlon, sus.th, sus.phi)
        keep = dict(th=_f64(sus.th), phi=_f64(sus.phi), w=_f64(sus.w), x=_f64(sus.x), a3=_f64(sus.a3),
                    f=_f64(sus.f), g=_f64(sus.g), detj=_f64(sus.detj), Acoef=_f64(sus.Acoef),
                    Bcoef=_f64(sus.Bcoef), area=_f64(sus.area), meshSize=_f64(sus.meshSize), spx=_f64(sus.spx),
                    spa3=_f64(sus.spa3), spdetj=_f64(sus.spdetj), spF=_f64(sus.spF), spG=_f64(sus.spG),
                    thG=thG, phiG=phiG, pw=pw)
        self._cells_keep = keep
        c = Cells()
        c.ncell, c.nlat, c.nlon = sus.ncell, sus.nlat, sus.nlon
        for k in ("th", "phi", "w", "x", "a3", "f", "g", "detj", "Acoef", "Bcoef", "area", "meshSize", "spx",
                  "spa3", "spdetj", "spF", "spG", "thG", "phiG"):
            setattr(c, k, _dp(keep[k]))
        c.patch_w = _dp(pw)
        c.patch_radius, c.nrad, c.nazm = radius, nrad, nazm
        self.cells = c
        self.patch = (radius, nrad, nazm, thG, phiG, pw)
        self.sus = sus
        return self

    def cell_targets(self, active=None):
        sus = self.sus
        n = sus.npoint
        npc = sus.nlat * sus.nlon
        p = np.arange(n)
        indx = np.stack([p // npc + 1, (p % npc) % sus.nlat + 1, (p % npc) // sus.nlat + 1]).astype(np.int32)
        A = np.repeat(sus.Acoef, npc)
        return self.make_targets(sus.x, A, indx, active)

    def make_targets(self, x, Acoef=None, indx=None, active=None):
        x = _f64(x)
        n = x.shape[1]
        A = np.full(n, 2.0) if Acoef is None else _f64(Acoef)          # TargetList_CreateFromRaw: Acoef = 2
        ix = np.full((3, n), -1, dtype=np.int32) if indx is None else np.ascontiguousarray(indx, dtype=np.int32)
        act = np.ones(n, dtype=np.int32) if active is None else np.ascontiguousarray(active, dtype=np.int32)
        t = Targets()
        t.n = n
        t.x, t.Acoef, t.indx, t.active = _dp(x), _dp(A), _ip(ix), _ip(act)
        t._keep = (x, A, ix, act)
        return t

    def add_int_on_rbcs(self, c1, c2, tl, v=None, flags=0):
        if v is None:
            v = np.zeros((3, tl.n))
        lib().orc_add_int_on_rbcs(C.byref(self.prm), C.byref(self.cells), C.c_double(c1), C.c_double(c2),
                                  C.byref(tl), _dp(v), C.c_int(flags))
        return v

    def sing_int(self, c1, c2, icell, ilat0, ilon0):
        dv = np.zeros(3)
        lib().orc_rbc_sing_int(C.byref(self.prm), C.byref(self.cells), C.c_double(c1), C.c_double(c2),
                               C.c_int(icell), C.c_int(ilat0), C.c_int(ilon0), _dp(dv))
        return dv

    def nearsing_int(self, c1, c2, icell, xi, x0, th0, phi0):
        dv = np.zeros(3)
        xi, x0 = _f64(xi), _f64(x0)
        lib().orc_rbc_nearsing_int(C.byref(self.prm), C.byref(self.cells), C.c_double(c1), C.c_double(c2),
                                   C.c_int(icell), _dp(xi), _dp(x0), C.c_double(th0), C.c_double(phi0), _dp(dv))
        return dv

    # ---- PME ----------------------------------------------------------------
    def pme(self):
        if self._pme is None:
            self._pme = C.c_void_p(lib().orc_pme_init(C.byref(self.prm)))
        return self._pme

    def pme_distrib(self, c1, c2, x, f=None, g=None, a3=None, Bcoef=None, accumulate=False):
        x, f, g, a3, Bcoef = _f64(x), _f64(f), _f64(g), _f64(a3), _f64(Bcoef)
        lib().orc_pme_distrib_source(self.pme(), C.c_double(c1), C.c_double(c2), C.c_int(x.shape[1]), _dp(x),
                                     _dp(f), _dp(g), _dp(a3), _dp(Bcoef), C.c_int(1 if accumulate else 0))

    def pme_transform(self):
        lib().orc_pme_transform(self.pme())

    def pme_interp(self, tl, v=None):
        if v is None:
            v = np.zeros((3, tl.n))
        lib().orc_pme_add_interp_vel(self.pme(), C.byref(tl), _dp(v))
        return v

    def pme_vv(self):
        Nx, Ny, Nz = self.Nb
        p = lib().orc_pme_vv(self.pme())
        return np.ctypeslib.as_array(p, shape=(3, Nz, Ny, Nx)).copy()

    def pme_bb(self):
        Nx, Ny, Nz = self.Nb
        p = lib().orc_pme_bb(self.pme())
        return np.ctypeslib.as_array(p, shape=(Nx // 2 + 1, Ny, Nz)).copy()

    def fft_forward(self, a):
        Nx, Ny, Nz = self.Nb
        a = _f64(a)
        out = np.zeros((Nz, Ny, Nx // 2 + 1), dtype=np.complex128)
        lib().orc_fft_forward(C.byref(self.prm), _dp(a), out.ctypes.data_as(c_dp))
        return out

    def fft_backward(self, a):
        Nx, Ny, Nz = self.Nb
        a = np.ascontiguousarray(a, dtype=np.complex128)
        out = np.zeros((Nz, Ny, Nx))
        lib().orc_fft_backward(C.byref(self.prm), a.ctypes.data_as(c_dp), _dp(out))
        return out

    def apply_cells(self, c1, c2, tl, v=None, flags=0, pme=True):
        if v is None:
            v = np.zeros((3, tl.n))
        lib().orc_apply_cells(C.byref(self.prm), C.byref(self.cells), self.pme() if pme else None,
                              C.c_double(c1), C.c_double(c2), C.byref(tl), _dp(v), C.c_int(flags))
        return v

    # ---- walls (rbc3d_oracle_walls.c) -----------------------------------------------
    def set_walls(self, W, ncell=None):
        """W: rbc3d_b200.synth.Walls.  Surface ids: cells 1..ncell, walls ncell+1.. (ModData: walls(1)%ID)."""
        if ncell is None:
            ncell = self.sus.ncell if self.cells is not None else 0
        keep = dict(nvert=np.ascontiguousarray(W.nvert, np.int32), nele=np.ascontiguousarray(W.nele, np.int32),
                    x=_f64(W.x), f=_f64(W.f), e2v=np.ascontiguousarray(W.e2v, np.int32), area=_f64(W.area),
                    epsDist=_f64(W.epsDist))
        w = WallsStruct()
        w.nwall, w.id0 = W.nwall, ncell + 1
        w.nvert, w.nele, w.e2v = _ip(keep["nvert"]), _ip(keep["nele"]), _ip(keep["e2v"])
        w.x, w.f, w.area, w.epsDist = _dp(keep["x"]), _dp(keep["f"]), _dp(keep["area"]), _dp(keep["epsDist"])
        self._walls_keep, self.walls, self.W = keep, w, W
        self.wall_mats = None
        return self

    def set_wall_traction(self, f):
        self._walls_keep["f"] = _f64(f).copy()
        self.walls.f = _dp(self._walls_keep["f"])

    def wall_targets(self, active=None):
        """tlist_wall: all wall vertices, Acoef = 2, indx(:,0) = wall id (ModTargetList.F90:122-131)."""
        W = self.W
        n = W.NV
        indx = np.full((3, n), -1, dtype=np.int32)
        indx[0] = self.walls.id0 + np.repeat(np.arange(W.nwall), W.nvert)
        return self.make_targets(W.x, np.full(n, 2.0), indx, active)

    def wall_compute_geometry(self):
        area, eps = np.zeros(self.W.NE), np.zeros(self.W.NE)
        lib().orc_wall_compute_geometry(C.byref(self.walls), _dp(area), _dp(eps))
        return area, eps

    def wall_centroids(self):
        xc = np.zeros((3, self.W.NE))
        lib().orc_wall_centroids(C.byref(self.prm), C.byref(self.walls), _dp(xc))
        return xc

    @staticmethod
    def min_dist_to_tri(xtar, xtri):
        """xtri: (3 corners, 3 comps).  Returns dist, s0, t0, x0."""
        xt, tri = _f64(xtar), _f64(xtri)
        s0, t0 = C.c_double(), C.c_double()
        x0 = np.zeros(3)
        d = lib().orc_min_dist_to_tri(_dp(xt), _dp(tri), C.byref(s0), C.byref(t0), _dp(x0))
        return d, s0.value, t0.value, x0

    def tri_int(self, xtri, ftri, xtar, s0=None, t0=None, want_lhs=True):
        """Tri_Int_Regular (s0 is None) or Tri_Int_Duffy.  Returns rhs(3), lhs(3 corners,3,3)."""
        tri, f, xt = _f64(xtri), _f64(ftri), _f64(xtar)
        rhs, lhs = np.zeros(3), np.zeros((3, 3, 3))
        if s0 is None:
            lib().orc_tri_int_regular(C.byref(self.prm), _dp(tri), _dp(f), _dp(xt), _dp(rhs), _dp(lhs) if want_lhs else None)
        else:
            lib().orc_tri_int_duffy(C.byref(self.prm), _dp(tri), _dp(f), _dp(xt), C.c_double(s0), C.c_double(t0),
                                    _dp(rhs), _dp(lhs) if want_lhs else None)
        return rhs, lhs

    def prepare_sing_int_on_walls(self, active=None):
        """PrepareSingIntOnWall for every wall; active: per wall-vertex flags of the wall target list or None."""
        self.free_wall_mats()
        vo = self.W.voff()
        mats = []
        for w in range(self.W.nwall):
            act = None if active is None else np.ascontiguousarray(active[vo[w]:vo[w + 1]], dtype=np.int32)
            mats.append(C.c_void_p(lib().orc_prepare_sing_int_on_wall(C.byref(self.prm), C.byref(self.walls), C.c_int(w), _ip(act))))
        self.wall_mats = mats
        return self

    def free_wall_mats(self):
        for m in getattr(self, "wall_mats", None) or []:
            lib().orc_wallmat_free(m)
        self.wall_mats = None

    def wall_matrix(self, iwall):
        """(rowptr, col, val[nblk,3,3]) of wall iwall's self-interaction matrix."""
        m = self.wall_mats[iwall]
        nblk = lib().orc_wallmat_nblk(m)
        rowptr = np.zeros(int(self.W.nvert[iwall]) + 1, np.int32)
        col = np.zeros(max(nblk, 1), np.int32)
        val = np.zeros((max(nblk, 1), 3, 3))
        lib().orc_wallmat_get(m, _ip(rowptr), _ip(col), _dp(val))
        return rowptr, col[:nblk], val[:nblk]

    def sing_int_on_wall(self, c1, iwall, f=None):
        vo = self.W.voff()
        fw = _f64(self._walls_keep["f"][:, vo[iwall]:vo[iwall + 1]] if f is None else f)
        v = np.zeros_like(fw)
        lib().orc_sing_int_on_wall(self.wall_mats[iwall], C.c_double(c1), _dp(fw), _dp(v))
        return v

    def add_int_on_walls(self, c1, tl, v=None):
        if v is None:
            v = np.zeros((3, tl.n))
        mats = None
        if self.wall_mats is not None:
            mats = (C.c_void_p * len(self.wall_mats))(*[m.value for m in self.wall_mats])
        lib().orc_add_int_on_walls(C.byref(self.prm), C.byref(self.walls), C.c_double(c1), C.byref(tl), _dp(v), mats)
        return v

    def wall_neighbor_signature(self, tl, self_skip=True):
        cnt = np.zeros(tl.n, np.int32)
        nd = np.zeros(tl.n, np.int32)
        sig = np.zeros(tl.n, np.uint64)
        lib().orc_wall_neighbor_signature(C.byref(self.prm), C.byref(self.walls), C.byref(tl), C.c_int(1 if self_skip else 0),
                                          _ip(cnt), sig.ctypes.data_as(C.POINTER(C.c_ulonglong)), _ip(nd))
        return cnt, sig, nd

    def pme_distrib_walls(self, c1, c2=0.0, accumulate=False):
        lib().orc_pme_distrib_walls(self.pme(), C.byref(self.prm), C.c_double(c1), C.c_double(c2), C.byref(self.walls),
                                    C.c_int(1 if accumulate else 0))

    # ---- ModRepulsion (closest-neighbour queries on the path's cell lists) ---------------------------
    def closest_neighbors(self, x, surf_id, eps_dist):
        """Closest_Neighbor_Cell / Closest_Neighbor_Wall for points x (3, n) lying on surfaces surf_id (n,):
        -> dist_cell, x0_cell, dist_wall, x0_wall (distances are inf where no neighbour is in the 27 list cells)."""
        x = _f64(x)
        n = x.shape[1]
        sid = np.ascontiguousarray(surf_id, dtype=np.int32)
        dc, dw = np.zeros(n), np.zeros(n)
        xc, xw = np.zeros((3, n)), np.zeros((3, n))
        W = C.byref(self.walls) if getattr(self, "walls", None) is not None else None
        lib().orc_closest_neighbors(C.byref(self.prm), C.byref(self.cells), W, C.c_int(n), _dp(x), _ip(sid),
                                    C.c_double(eps_dist), _dp(dc), _dp(xc), _dp(dw), _dp(xw))
        return dc, xc, dw, xw

    def inter_cell_repulsion(self, eps_dist, active=None):
        """InterCellRepulsion up to the displacement: -> dx (3, Np), number of moved points, min separation."""
        n = self.sus.npoint
        dx = np.zeros((3, n))
        dmin = C.c_double()
        act = None if active is None else np.ascontiguousarray(active, dtype=np.int32)
        W = C.byref(self.walls) if getattr(self, "walls", None) is not None else None
        lib().orc_inter_cell_repulsion.restype = C.c_int
        cnt = lib().orc_inter_cell_repulsion(C.byref(self.prm), C.byref(self.cells), W, _ip(act), C.c_double(eps_dist),
                                             _dp(dx), C.byref(dmin))
        return dx, int(cnt), dmin.value

    def apply(self, c1, c2, tl, cells=True, walls=False, v=None):
        """v += AddIntOnRbcs + AddIntOnWalls + PME over the chosen source sets (the composition of
        ModVelSolver.F90:473-493 / ModNoSlip.F90:172-191, 284-299)."""
        if v is None:
            v = np.zeros((3, tl.n))
        if cells:
            self.add_int_on_rbcs(c1, c2, tl, v)
        if walls:
            self.add_int_on_walls(c1, tl, v)
        first = True
        if cells:
            sus = self.sus
            npc = sus.nlat * sus.nlon
            self.pme_distrib(c1, c2, sus.x, sus.weighted(sus.f) if abs(c1) > 1e-10 else None,
                             sus.weighted(sus.g) if abs(c2) > 1e-10 else None, sus.a3, np.repeat(sus.Bcoef, npc))
            first = False
        if walls:
            self.pme_distrib_walls(c1, c2, accumulate=not first)
        self.pme_transform()
        self.pme_interp(tl, v)
        return v

    def __del__(self):
        try:
            self.free_wall_mats()
        except Exception:
            pass
        try:
            if self._pme is not None:
                lib().orc_pme_finalize(self._pme)
        except Exception:
            pass
