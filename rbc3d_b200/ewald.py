"""Host-side mirror of the reference's operator interface for the Ewald boundary-integral path.

The method names and the call protocol are the reference's module procedures:

    ModConf      SetEwaldPrms
    ModEwaldFunc EwaldCoeff_SL / EwaldCoeff_DL / EwaldCoeff_SL_Exact / EwaldCoeff_DL_Exact
    ModSourceList / ModTargetList   SourceList_UpdateCoord, SourceList_UpdateDensity, TargetList_CreateFromRaw
    ModIntOnRbcs AddIntOnRbcs(c1, c2, tlist, v)
    ModPME       PME_Distrib_Source(c1, c2, cells, walls) -> PME_Transform() -> PME_Add_Interp_Vel(tlist, v)

plus the fused ``apply`` the GMRES callbacks amount to (ModVelSolver.F90:568-584).  Everything forwards to
librbc3d_b200.so through the C ABI (rbc3d_b200.capi); v arrays are SoA (3, n) NumPy arrays, accumulated into like
the Fortran ``v(:, :)`` arguments.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import TL_CELLS, TL_RAW, TL_WALLS, check, dp, f64, i32, ip


def SetEwaldPrms(Lb, alpha=0.44, eps=1e-3, P=8, nranks=1):
    """ModConf.F90:348-408 -> (rc, Nb)."""
    lib = capi.load()
    Lb3 = (C.c_double * 3)(*[float(v) for v in Lb])
    rc = C.c_double()
    Nb = (C.c_int32 * 3)()
    check(lib.rbc3d_set_ewald_prms(Lb3, alpha, eps, P, nranks, C.byref(rc), Nb), "rbc3d_set_ewald_prms")
    return rc.value, [int(v) for v in Nb]


def EwaldCoeff_SL_Exact(r, alpha):
    a, b = C.c_double(), C.c_double()
    check(capi.load().rbc3d_ewald_coeff_sl_exact(r, alpha, C.byref(a), C.byref(b)))
    return a.value, b.value


def EwaldCoeff_DL_Exact(r, alpha):
    a = C.c_double()
    check(capi.load().rbc3d_ewald_coeff_dl_exact(r, alpha, C.byref(a)))
    return a.value


class EwaldOperator:
    """One context per GPU: the module state of ModConf / ModData / ModPME behind the drop-in boundary."""

    def __init__(self, Lb, alpha=0.44, eps=1e-3, P=8, rc=None, Nb=None, device=-1, nranks=1):
        self.lib = capi.load()
        self.Lb = np.asarray(Lb, dtype=np.float64)
        self.alpha, self.eps, self.P = float(alpha), float(eps), int(P)
        rc0, Nb0 = SetEwaldPrms(self.Lb, alpha, eps, P, nranks)
        self.rc = rc0 if rc is None else float(rc)
        self.Nb = Nb0 if Nb is None else [int(v) for v in Nb]
        self._h = C.c_void_p()
        Lb3 = (C.c_double * 3)(*self.Lb)
        Nb3 = (C.c_int32 * 3)(*self.Nb)
        check(self.lib.rbc3d_ctx_create(C.byref(self._h), Lb3, self.alpha, self.eps, self.P, self.rc, Nb3, device),
              "rbc3d_ctx_create (PME_Init)")
        self.ncell = self.npoint = 0
        self.nranks, self.rank = 1, 0
        self.n_raw = 0
        self._keep = {}

    # -- lifetime ---------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.rbc3d_ctx_destroy(self._h)  # PME_Finalize
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- multi-GPU --------------------------------------------------------------------------------------
    def attach_comm(self, nranks, rank, dist=None, unique_id: bytes | None = None):
        """Join the NCCL communicator of the run (before any geometry is set).  The 128-byte unique id is made on
        rank 0 and handed round by the host program: torch.distributed here, MPI_Bcast in the Fortran driver."""
        self.nranks, self.rank = int(nranks), int(rank)
        if nranks == 1:
            return
        if unique_id is None:
            buf = C.create_string_buffer(128)
            if rank == 0:
                check(self.lib.rbc3d_comm_unique_id(buf), "rbc3d_comm_unique_id")
            obj = [bytes(buf.raw) if rank == 0 else None]
            dist.broadcast_object_list(obj, src=0)
            unique_id = obj[0]
        idbuf = C.create_string_buffer(unique_id, 128)
        check(self.lib.rbc3d_ctx_attach_comm(self._h, nranks, rank, idbuf), "rbc3d_ctx_attach_comm")

    def ownership_mask(self, sus, nranks, rank, by="zslab"):
        """active flags of the targets this rank computes: whole cells, by the z-slab of their centroid (the slabs of the
        PME mesh, DomainDecomp) or -- by="block" -- by contiguous blocks of cell indices"""
        from . import partition
        if by == "block":
            return partition.ownership_mask(sus.ncell, sus.nlat * sus.nlon, nranks, rank)
        return partition.ownership_mask_zslab(sus.x, sus.nlat * sus.nlon, self.Lb, nranks, rank)

    def TargetList_CollectArray(self, v, tlist=TL_CELLS):
        """v <- sum over ranks (ModTargetList.F90:172-202)."""
        check(self.lib.rbc3d_collect_array(self._h, tlist, dp(v)), "TargetList_CollectArray")
        return v

    # -- ModEwaldFunc -----------------------------------------------------------------------------------
    def EwaldCoeff_SL(self, r):
        a, b = C.c_double(), C.c_double()
        check(self.lib.rbc3d_ewald_coeff_sl(self._h, r, C.byref(a), C.byref(b)))
        return a.value, b.value

    def EwaldCoeff_DL(self, r):
        a = C.c_double()
        check(self.lib.rbc3d_ewald_coeff_dl(self._h, r, C.byref(a)))
        return a.value

    # -- cells ------------------------------------------------------------------------------------------
    def set_mesh(self, ncell, nlat, nlon, th, phi, w):
        th, phi, w = f64(th), f64(phi), f64(w)
        check(self.lib.rbc3d_cells_set_mesh(self._h, ncell, nlat, nlon, dp(th), dp(phi), dp(w)), "rbc3d_cells_set_mesh")
        self.ncell, self.nlat, self.nlon = ncell, nlat, nlon
        self.npoint = ncell * nlat * nlon
        # shared polar patch size, RbcPolarPatch_Create (ModPolarPatch.F90:42-45)
        self.nrad = 2 * int(round((np.pi / np.sqrt(nlat)) / (np.pi / nlat)))
        self.nazm = 2 * self.nrad

    def SourceList_UpdateCoord(self, x, a3, Acoef, Bcoef, area, meshSize, spx, spa3, spdetj, active=None):
        """SourceList_UpdateCoord + TargetList_Update for the cell lists (rebuilds the cell list)."""
        arrs = [f64(a) for a in (x, a3, Acoef, Bcoef, area, meshSize, spx, spa3, spdetj)]
        act = i32(active)
        check(self.lib.rbc3d_cells_set_geometry(self._h, *[dp(a) for a in arrs], ip(act)), "rbc3d_cells_set_geometry")

    def SourceList_UpdateCoord_mesh(self, x, a3, detj, Acoef, Bcoef, area, meshSize, active=None):
        """SourceList_UpdateCoord with the geometry splines (Rbc_BuildSurfaceSource(xFlag)) built on the device from
        the mesh fields; needs enable_device_splines."""
        arrs = [f64(a) for a in (x, a3, detj, Acoef, Bcoef, area, meshSize)]
        act = i32(active)
        check(self.lib.rbc3d_cells_set_geometry_mesh(self._h, *[dp(a) for a in arrs], ip(act)),
              "rbc3d_cells_set_geometry_mesh")

    def get_geometry_spline(self, which="x"):
        """device copy of spline(x) / spline(a3) / spline(detJ) in the ABI layout (tests)."""
        nv = 1 if which == "detj" else 3
        out = np.zeros((self.ncell, 4, nv, self.nlon, 2 * self.nlat))
        check(self.lib.rbc3d_cells_get_geometry_spline(self._h, {"x": 0, "a3": 1, "detj": 2}[which], dp(out)),
              "rbc3d_cells_get_geometry_spline")
        return out

    def SourceList_UpdateDensity(self, f=None, g=None, spF=None, spG=None):
        """f, g: slist%f / slist%g (densities * detJ * w); spF, spG: splines of f*detJ, g*detJ."""
        f, g, spF, spG = f64(f), f64(g), f64(spF), f64(spG)
        check(self.lib.rbc3d_cells_set_density(self._h, dp(f), dp(g), dp(spF), dp(spG)), "rbc3d_cells_set_density")

    def enable_device_splines(self, nlat0):
        """Rbc_BuildSurfaceSource(fFlag/gFlag) on the GPU: densities passed without splines get them built there."""
        check(self.lib.rbc3d_cells_enable_device_splines(self._h, int(nlat0)), "rbc3d_cells_enable_device_splines")

    def get_density_spline(self, which="g"):
        """device copy of spline(f detJ) / spline(g detJ) in the ABI layout (tests)."""
        out = np.zeros((self.ncell, 4, 3, self.nlon, 2 * self.nlat))
        check(self.lib.rbc3d_cells_get_density_spline(self._h, 1 if which == "g" else 0, dp(out)), "get_density_spline")
        return out

    def set_suspension(self, sus, active=None, with_f=True, with_g=True):
        """Convenience: load a rbc3d_b200.synth.Suspension."""
        self.set_mesh(sus.ncell, sus.nlat, sus.nlon, sus.th, sus.phi, sus.w)
        self.SourceList_UpdateCoord(sus.x, sus.a3, sus.Acoef, sus.Bcoef, sus.area, sus.meshSize, sus.spx, sus.spa3,
                                    sus.spdetj, active)
        f = sus.weighted(sus.f) if (with_f and sus.f is not None) else None
        g = sus.weighted(sus.g) if (with_g and sus.g is not None) else None
        self.SourceList_UpdateDensity(f, g, sus.spF if f is not None else None, sus.spG if g is not None else None)

    def TargetList_CreateFromRaw(self, x, active=None):
        x = f64(x)
        act = i32(active)
        self.n_raw = x.shape[1]
        check(self.lib.rbc3d_targets_set_raw(self._h, self.n_raw, dp(x), ip(act)), "rbc3d_targets_set_raw")

    def _n(self, tlist):
        return self.npoint if tlist == TL_CELLS else (self.n_raw if tlist == TL_RAW else self.n_wall_vert)

    # -- walls (ModIntOnWalls, wall branches of ModSourceList / ModTargetList) --------------------------------
    def set_walls(self, W, active=None, traction=True):
        """W: rbc3d_b200.synth.Walls (or any object with nvert, nele, x, e2v, area, epsDist[, f]).
        SourceList_UpdateCoord(slist_wall) + TargetList_Update(tlist_wall)."""
        nv, ne = i32(W.nvert), i32(W.nele)
        x, e2v, area, eps = f64(W.x), i32(W.e2v), f64(W.area), f64(W.epsDist)
        act = i32(active)
        self.n_wall_vert, self.n_wall_ele, self.wall_nvert = int(nv.sum()), int(ne.sum()), [int(v) for v in nv]
        check(self.lib.rbc3d_walls_set(self._h, len(nv), ip(nv), ip(ne), dp(x), ip(e2v), dp(area), dp(eps), ip(act)),
              "rbc3d_walls_set")
        if traction and getattr(W, "f", None) is not None:
            self.set_wall_traction(W.f)

    def set_wall_traction(self, f):
        f = f64(f)
        assert f.shape == (3, self.n_wall_vert)
        check(self.lib.rbc3d_walls_set_traction(self._h, dp(f)), "rbc3d_walls_set_traction")

    def PrepareSingIntOnWall(self):
        """All walls at once (the reference loops over walls, ModTimeInt / ModNoSlip callers)."""
        check(self.lib.rbc3d_wall_prepare_sing(self._h), "PrepareSingIntOnWall")

    def SingIntOnWall(self, c1, iwall):
        v = np.zeros((3, self.wall_nvert[iwall]))
        check(self.lib.rbc3d_sing_int_on_wall(self._h, c1, iwall, dp(v)), "SingIntOnWall")
        return v

    def AddIntOnWalls(self, c1, tlist=TL_CELLS, v=None):
        v = self._v(tlist, v)
        check(self.lib.rbc3d_add_int_on_walls(self._h, c1, tlist, dp(v)), "AddIntOnWalls")
        return v

    def MinDistToTri(self, xtar, xtri):
        """xtar (3, n) SoA, xtri (n, 3 corners, 3).  Returns dist, s0, t0."""
        xtar, xtri = f64(xtar), f64(xtri)
        n = xtar.shape[1]
        d, s0, t0 = np.zeros(n), np.zeros(n), np.zeros(n)
        check(self.lib.rbc3d_min_dist_to_tri(self._h, n, dp(xtar), dp(xtri), dp(d), dp(s0), dp(t0)), "MinDistToTri")
        return d, s0, t0

    def Tri_Int(self, xtri, ftri, xtar, s0=None, t0=None, lhs=False):
        """Tri_Int_Regular (s0 is None) / Tri_Int_Duffy for n triples: xtri, ftri (n,3,3); xtar (3,n) SoA.
        Returns rhs (n,3) [, lhs (n,3,3,3)]."""
        xtri, ftri, xtar, s0, t0 = f64(xtri), f64(ftri), f64(xtar), f64(s0), f64(t0)
        n = xtar.shape[1]
        rhs = np.zeros((n, 3))
        L = np.zeros((n, 3, 3, 3)) if lhs else None
        check(self.lib.rbc3d_tri_int(self._h, n, dp(xtri), dp(ftri), dp(xtar), dp(s0), dp(t0), dp(rhs), dp(L)), "Tri_Int")
        return (rhs, L) if lhs else rhs

    def wall_matrix(self):
        nb = (C.c_int32 * 1)()
        rowptr = np.zeros(self.n_wall_vert + 1, np.int32)
        check(self.lib.rbc3d_wall_matrix_get(self._h, nb, ip(rowptr), None, None, 0))
        col = np.zeros(max(nb[0], 1), np.int32)
        val = np.zeros((max(nb[0], 1), 3, 3))
        check(self.lib.rbc3d_wall_matrix_get(self._h, nb, ip(rowptr), ip(col), dp(val), nb[0]))
        return rowptr, col[:nb[0]], val[:nb[0]]

    def wall_neighbor_signature(self, tlist=TL_CELLS, self_skip=True):
        n = self._n(tlist)
        cnt, nd, sig = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.uint64)
        check(self.lib.rbc3d_wall_neighbor_signature(self._h, tlist, int(self_skip), ip(cnt), sig.ctypes.data_as(capi.c_up),
                                                     ip(nd)))
        return cnt, sig, nd

    def _v(self, tlist, v):
        if v is None:
            return np.zeros((3, self._n(tlist)))
        assert v.dtype == np.float64 and v.flags.c_contiguous and v.shape == (3, self._n(tlist))
        return v

    # -- operator pieces (reference names) ----------------------------------------------------------------
    def AddIntOnRbcs(self, c1, c2, tlist=TL_CELLS, v=None):
        v = self._v(tlist, v)
        check(self.lib.rbc3d_add_int_on_rbcs(self._h, c1, c2, tlist, dp(v)), "AddIntOnRbcs")
        return v

    def PME_Distrib_Source(self, c1, c2, cells=True, walls=False):
        check(self.lib.rbc3d_pme_distrib_source(self._h, c1, c2, int(cells), int(walls)), "PME_Distrib_Source")

    def PME_Transform(self):
        check(self.lib.rbc3d_pme_transform(self._h), "PME_Transform")

    def PME_Add_Interp_Vel(self, tlist=TL_CELLS, v=None):
        v = self._v(tlist, v)
        check(self.lib.rbc3d_pme_add_interp_vel(self._h, tlist, dp(v)), "PME_Add_Interp_Vel")
        return v

    def apply(self, c1, c2, tlist=TL_CELLS, v=None, cells=True, walls=False):
        v = self._v(tlist, v)
        check(self.lib.rbc3d_apply(self._h, c1, c2, int(cells), int(walls), tlist, dp(v)), "rbc3d_apply")
        return v

    def apply_assign(self, c1, c2, tlist=TL_CELLS, v=None, cells=True, walls=False):
        """v = operator (the caller's v = 0 folded in; v is not uploaded)."""
        v = self._v(tlist, v)
        check(self.lib.rbc3d_apply_assign(self._h, c1, c2, int(cells), int(walls), tlist, dp(v)), "rbc3d_apply_assign")
        return v

    def apply_collect(self, c1, c2, tlist=TL_CELLS, v=None, cells=True, walls=False):
        """v = operator summed over the ranks (apply_assign + TargetList_CollectArray, reduced on the devices)."""
        v = self._v(tlist, v)
        check(self.lib.rbc3d_apply_collect(self._h, c1, c2, int(cells), int(walls), tlist, dp(v)), "rbc3d_apply_collect")
        return v

    def apply_resident(self, c1, c2, tlist=TL_CELLS, cells=True, walls=False):
        check(self.lib.rbc3d_apply_resident(self._h, c1, c2, int(cells), int(walls), tlist), "rbc3d_apply_resident")

    def get_velocity(self, tlist=TL_CELLS):
        v = np.zeros((3, self._n(tlist)))
        check(self.lib.rbc3d_get_velocity(self._h, tlist, dp(v)), "rbc3d_get_velocity")
        return v

    def set_skip_flags(self, flags):
        check(self.lib.rbc3d_set_skip_flags(self._h, flags))

    def set_pair_self(self, mode):
        check(self.lib.rbc3d_set_pair_self(self._h, int(mode)))

    # -- cell velocity solve on the device --------------------------------------------------
    def solver_setup(self, nlat0, detj):
        check(self.lib.rbc3d_solver_setup(self._h, int(nlat0), dp(f64(detj))), "rbc3d_solver_setup")
        n = C.c_int64()
        check(self.lib.rbc3d_solver_dof(self._h, C.byref(n)))
        self.solver_dof = int(n.value)              # unknowns THIS rank holds (all of them on one rank)
        m = C.c_int32()
        cells = np.zeros(max(self.ncell, 1), np.int32)
        check(self.lib.rbc3d_solver_cells(self._h, C.byref(m), ip(cells), self.ncell))
        self.solver_cells = cells[:m.value].copy()  # ... of these cells, in vector order

    def solver_matmult(self, u, out=None):
        """MyMatMult on the device: packed SH coefficients in, packed SH coefficients out (of this rank's cells)."""
        u = f64(u)
        b = np.zeros(self.solver_dof) if out is None else out
        check(self.lib.rbc3d_solver_matmult(self._h, dp(u), dp(b)), "rbc3d_solver_matmult")
        return b

    def solver_rhs(self, vbkg=(1.0, 0.0, 0.0), walls=False):
        """Compute_Rhs on the device (needs the single-layer density set): packed coefficients."""
        rhs = np.zeros(self.solver_dof)
        vb = f64(np.asarray(vbkg, dtype=np.float64))
        check(self.lib.rbc3d_solver_rhs(self._h, dp(vb), int(walls), dp(rhs)), "rbc3d_solver_rhs")
        return rhs

    def solver_velocity(self, sol):
        v = np.zeros((3, self.npoint))
        check(self.lib.rbc3d_solver_velocity(self._h, dp(f64(sol)), dp(v)), "rbc3d_solver_velocity")
        return v

    def solver_gmres(self, rhs, x0=None, rtol=1e-11, restart=30, maxit=200):
        """-> (sol, niter, residual history)."""
        rhs = f64(rhs)
        sol = np.zeros(self.solver_dof) if x0 is None else np.array(x0, dtype=np.float64)
        hist = np.full(maxit + 1, np.nan)
        nit = C.c_int()
        check(self.lib.rbc3d_solver_gmres(self._h, dp(rhs), dp(sol), float(rtol), int(restart), int(maxit), C.byref(nit),
                                          dp(hist)), "rbc3d_solver_gmres")
        return sol, nit.value, hist[:nit.value + 1]

    # -- NoSlipWall on the device (ModNoSlip.F90:44-149) -------------------------------------------------------
    def noslip_solve(self, f, indx, nindep, vbkg, cells=True, rtol=1e-3, maxit=60, want_slip=True):
        """-> (f_new (3, NV), niter, residual history, residual wall velocity (3, NV) or None)."""
        f = np.array(f, dtype=np.float64, order="C")
        indx = i32(indx)
        vb = f64(np.asarray(vbkg, dtype=np.float64))
        hist = np.full(maxit + 1, np.nan)
        slip = np.zeros_like(f) if want_slip else None
        nit = C.c_int()
        check(self.lib.rbc3d_noslip_solve(self._h, ip(indx), int(nindep), dp(vb), int(bool(cells)), float(rtol), int(maxit),
                                          dp(f), C.byref(nit), dp(hist), dp(slip)), "rbc3d_noslip_solve")
        return f, nit.value, hist[:nit.value + 1], slip

    # -- ModRepulsion closest-neighbour queries on the GPU cell lists ------------------------
    def closest_neighbors(self, x, surf_id, eps_dist):
        """Closest_Neighbor_Cell / Closest_Neighbor_Wall for points x (3, n) on surfaces surf_id (n,) ->
        dist_cell, x0_cell, dist_wall, x0_wall (inf where no other surface is in the 27 neighbouring list cells)."""
        x = f64(x)
        n = x.shape[1]
        sid = i32(surf_id)
        dc, dw = np.zeros(n), np.zeros(n)
        xc, xw = np.zeros((3, n)), np.zeros((3, n))
        check(self.lib.rbc3d_closest_neighbors(self._h, n, dp(x), ip(sid), float(eps_dist), dp(dc), dp(xc), dp(dw), dp(xw)),
              "rbc3d_closest_neighbors")
        return dc, xc, dw, xw

    def sing_cache_info(self):
        """(cached path active, patch points per target streamed from the cache)."""
        a, b = C.c_int32(), C.c_int32()
        check(self.lib.rbc3d_sing_cache_info(self._h, C.byref(a), C.byref(b)))
        return bool(a.value), b.value

    def set_replicated_density(self, on=True):
        """host densities identical on all ranks: upload 1/nranks each, all-gather over NVLink (collective)."""
        check(self.lib.rbc3d_set_replicated_density(self._h, int(bool(on))))

    def set_overlap(self, mode):
        """-1 auto, 0 one stream, 1 PME chain on its own stream beside the real-space kernels (rbc3d_set_overlap)"""
        check(self.lib.rbc3d_set_overlap(self._h, int(mode)), "rbc3d_set_overlap")

    def pair_cache_info(self):
        """(cells cached, 256-byte coefficient rows) of the same-surface double-layer pair cache."""
        nc, rows = C.c_int32(), C.c_int64()
        check(self.lib.rbc3d_pair_cache_info(self._h, C.byref(nc), C.byref(rows)))
        return nc.value, rows.value

    def set_sing_cache(self, mode):
        check(self.lib.rbc3d_set_sing_cache(self._h, int(mode)))

    # -- introspection ------------------------------------------------------------------------------------
    def cell_list(self):
        Nc = (C.c_int32 * 3)()
        check(self.lib.rbc3d_cell_list_get(self._h, Nc, None, None, None))
        ncells = Nc[0] * Nc[1] * Nc[2]
        cid = np.zeros(self.npoint, dtype=np.int32)
        order = np.zeros(self.npoint, dtype=np.int32)
        start = np.zeros(ncells + 1, dtype=np.int32)
        check(self.lib.rbc3d_cell_list_get(self._h, Nc, ip(cid), ip(order), ip(start)))
        return [int(v) for v in Nc], cid, order, start

    def cell_list_dims(self):
        Nc = (C.c_int32 * 3)()
        check(self.lib.rbc3d_cell_list_get(self._h, Nc, None, None, None))
        return [int(v) for v in Nc]

    def neighbor_signature(self, tlist=TL_CELLS):
        n = self._n(tlist)
        cnt = np.zeros(n, dtype=np.int32)
        sig = np.zeros(n, dtype=np.uint64)
        check(self.lib.rbc3d_neighbor_signature(self._h, tlist, ip(cnt), sig.ctypes.data_as(capi.c_up)))
        return cnt, sig

    def nearsing_entries(self, tlist=TL_CELLS):
        n = C.c_int()
        check(self.lib.rbc3d_nearsing_get(self._h, tlist, C.byref(n), None, None, None, None, None, None, 0))
        m = n.value
        out = dict(target=np.zeros(m, np.int32), cell=np.zeros(m, np.int32), flag=np.zeros(m, np.int32),
                   th0=np.zeros(m), phi0=np.zeros(m), dist=np.zeros(m))
        if m:
            check(self.lib.rbc3d_nearsing_get(self._h, tlist, C.byref(n), ip(out["target"]), ip(out["cell"]),
                                              ip(out["flag"]), dp(out["th0"]), dp(out["phi0"]), dp(out["dist"]), m))
        return out

    def pme_grid(self):
        Nx, Ny, Nz = self.Nb
        vv = np.zeros((3, Nz, Ny, Nx))
        check(self.lib.rbc3d_pme_get_grid(self._h, dp(vv)))
        return vv

    def timings(self):
        ms = (C.c_float * len(capi.STAGES))()
        check(self.lib.rbc3d_get_timings(self._h, ms))
        return {k: float(ms[i]) for i, k in enumerate(capi.STAGES)}

    def launch_count(self):
        n = C.c_longlong()
        check(self.lib.rbc3d_get_launch_count(self._h, C.byref(n)))
        return n.value


def measure_fp64_peak(device=-1):
    t = C.c_double()
    check(capi.load().rbc3d_measure_fp64_peak(device, C.byref(t)), "rbc3d_measure_fp64_peak")
    return t.value
