"""ctypes binding of librbc3d_b200.so (include/rbc3d.h).  Plain pointers and sizes only; NumPy arrays are the
host buffers.  The library is the only compute path: if it is missing or no CUDA device is present every operator
call raises -- there is no CPU fallback in the product (the oracle under oracle/ is test infrastructure)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librbc3d_b200.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_up = C.POINTER(C.c_uint64)
c_fp = C.POINTER(C.c_float)

TL_CELLS, TL_RAW, TL_WALLS = 0, 1, 2
STAGES = ["pair", "sing", "nearsing", "linear", "spread", "fft", "kspace", "fft_inv", "interp", "combine", "wall",
          "comm", "h2d", "d2h", "density", "total", "pme_chain", "real_chain"]

# name -> (restype, argtypes); every symbol declared in include/rbc3d.h
SIGNATURES = {
    "rbc3d_last_error": (C.c_char_p, []),
    "rbc3d_version": (C.c_int, []),
    "rbc3d_set_ewald_prms": (C.c_int, [c_dp, C.c_double, C.c_double, C.c_int, C.c_int, c_dp, c_ip]),
    "rbc3d_ctx_create": (C.c_int, [C.POINTER(C.c_void_p), c_dp, C.c_double, C.c_double, C.c_int, C.c_double, c_ip,
                                   C.c_int]),
    "rbc3d_ctx_destroy": (C.c_int, [C.c_void_p]),
    "rbc3d_comm_unique_id": (C.c_int, [C.c_void_p]),
    "rbc3d_ctx_attach_comm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "rbc3d_collect_array": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "rbc3d_ewald_coeff_sl_exact": (C.c_int, [C.c_double, C.c_double, c_dp, c_dp]),
    "rbc3d_ewald_coeff_dl_exact": (C.c_int, [C.c_double, C.c_double, c_dp]),
    "rbc3d_ewald_coeff_sl": (C.c_int, [C.c_void_p, C.c_double, c_dp, c_dp]),
    "rbc3d_ewald_coeff_dl": (C.c_int, [C.c_void_p, C.c_double, c_dp]),
    "rbc3d_cells_set_mesh": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp]),
    "rbc3d_cells_set_geometry": (C.c_int, [C.c_void_p] + [c_dp] * 9 + [c_ip]),
    "rbc3d_cells_set_geometry_mesh": (C.c_int, [C.c_void_p] + [c_dp] * 7 + [c_ip]),
    "rbc3d_cells_set_density": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp, c_dp]),
    "rbc3d_cells_enable_device_splines": (C.c_int, [C.c_void_p, C.c_int]),
    "rbc3d_cells_get_density_spline": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "rbc3d_cells_get_geometry_spline": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "rbc3d_walls_set": (C.c_int, [C.c_void_p, C.c_int, c_ip, c_ip, c_dp, c_ip, c_dp, c_dp, c_ip]),
    "rbc3d_walls_set_traction": (C.c_int, [C.c_void_p, c_dp]),
    "rbc3d_wall_prepare_sing": (C.c_int, [C.c_void_p]),
    "rbc3d_sing_int_on_wall": (C.c_int, [C.c_void_p, C.c_double, C.c_int, c_dp]),
    "rbc3d_add_int_on_walls": (C.c_int, [C.c_void_p, C.c_double, C.c_int, c_dp]),
    "rbc3d_min_dist_to_tri": (C.c_int, [C.c_void_p, C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "rbc3d_tri_int": (C.c_int, [C.c_void_p, C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "rbc3d_wall_matrix_get": (C.c_int, [C.c_void_p, c_ip, c_ip, c_ip, c_dp, C.c_int]),
    "rbc3d_wall_neighbor_signature": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_ip, c_up, c_ip]),
    "rbc3d_targets_set_raw": (C.c_int, [C.c_void_p, C.c_int, c_dp, c_ip]),
    "rbc3d_add_int_on_rbcs": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int, c_dp]),
    "rbc3d_pme_distrib_source": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int]),
    "rbc3d_pme_transform": (C.c_int, [C.c_void_p]),
    "rbc3d_pme_add_interp_vel": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "rbc3d_apply": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, c_dp]),
    "rbc3d_apply_assign": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, c_dp]),
    "rbc3d_apply_collect": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, c_dp]),
    "rbc3d_apply_resident": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int]),
    "rbc3d_get_velocity": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "rbc3d_set_skip_flags": (C.c_int, [C.c_void_p, C.c_int]),
    "rbc3d_set_sing_cache": (C.c_int, [C.c_void_p, C.c_int]),
    "rbc3d_set_pair_self": (C.c_int, [C.c_void_p, C.c_int]),
    "rbc3d_solver_setup": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "rbc3d_solver_dof": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "rbc3d_noslip_solve": (C.c_int, [C.c_void_p, c_ip, C.c_int, c_dp, C.c_int, C.c_double, C.c_int, c_dp,
                                    C.POINTER(C.c_int), c_dp, c_dp]),
    "rbc3d_closest_neighbors": (C.c_int, [C.c_void_p, C.c_int, c_dp, c_ip, C.c_double, c_dp, c_dp, c_dp, c_dp]),
    "rbc3d_solver_cells": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int]),
    "rbc3d_solver_matmult": (C.c_int, [C.c_void_p, c_dp, c_dp]),
    "rbc3d_solver_rhs": (C.c_int, [C.c_void_p, c_dp, C.c_int, c_dp]),
    "rbc3d_solver_velocity": (C.c_int, [C.c_void_p, c_dp, c_dp]),
    "rbc3d_solver_gmres": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_int), c_dp]),
    "rbc3d_sing_cache_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "rbc3d_set_replicated_density": (C.c_int, [C.c_void_p, C.c_int]),
    "rbc3d_set_overlap": (C.c_int, [C.c_void_p, C.c_int]),
    "rbc3d_pair_cache_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "rbc3d_cell_list_get": (C.c_int, [C.c_void_p, c_ip, c_ip, c_ip, c_ip]),
    "rbc3d_neighbor_signature": (C.c_int, [C.c_void_p, C.c_int, c_ip, c_up]),
    "rbc3d_nearsing_get": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), c_ip, c_ip, c_ip, c_dp, c_dp, c_dp,
                                     C.c_int]),
    "rbc3d_pme_get_grid": (C.c_int, [C.c_void_p, c_dp]),
    "rbc3d_get_timings": (C.c_int, [C.c_void_p, c_fp]),
    "rbc3d_get_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_longlong)]),
    "rbc3d_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "rbc3d_host_unregister": (C.c_int, [C.c_void_p]),
    "rbc3d_measure_fp64_peak": (C.c_int, [C.c_int, c_dp]),
}

_lib = None


class Rbc3dError(RuntimeError):
    pass


def load():
    """Load the shared library (raises if it was not built: the product has no other path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Rbc3dError(f"{LIB_PATH} not built; run `python -m rbc3d_b200.build` (needs nvcc). "
                             "There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().rbc3d_last_error().decode(errors="replace")
        raise Rbc3dError(f"{what} failed with code {rc}: {msg}")


def dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def ip(a):
    return None if a is None else a.ctypes.data_as(c_ip)


def f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)
