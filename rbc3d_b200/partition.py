"""Host-side partition logic of the multi-GPU operator (pure NumPy; shared by bench.py, the tests and the harness).

The reference assigns TARGETS to MPI ranks by z-slab, point by point (SetActiveFlag, ModTargetList.F90:205-233:
``floor(z * iLb(3) * nranks) mod nranks``) and replicates all sources.  The B200 path keeps that contract -- the
caller passes ``active`` flags and sums the per-rank results (TargetList_CollectArray) -- but the harness assigns
whole cells to ranks (contiguous index blocks, the same blocks whose sources a rank spreads on the PME mesh), so
that a cell's singular integrals and spline data are needed on one GPU only.
"""
from __future__ import annotations

import numpy as np


def cell_block(ncell: int, nranks: int, rank: int) -> tuple[int, int]:
    """[lo, hi) of the cells owned by ``rank`` -- must match rbc3d_cells_set_geometry (capi.cu: c_lo, c_hi)."""
    return (ncell * rank) // nranks, (ncell * (rank + 1)) // nranks


def ownership_mask(ncell: int, npc: int, nranks: int, rank: int) -> np.ndarray:
    """active flags (int32, one per cell point) of the targets a rank computes."""
    lo, hi = cell_block(ncell, nranks, rank)
    act = np.zeros(ncell * npc, dtype=np.int32)
    act[lo * npc:hi * npc] = 1
    return act


def zslab_active(x: np.ndarray, Lb, nranks: int, rank: int) -> np.ndarray:
    """The reference's point-wise z-slab ownership (SetActiveFlag, ModTargetList.F90:221-222)."""
    iz = np.floor(x[2] * (1.0 / Lb[2]) * nranks).astype(np.int64)
    return (np.mod(iz, nranks) == rank).astype(np.int32)
