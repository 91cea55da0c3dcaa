"""Generate tests/golden/oracle_8cells.npz: velocities of the two cell operators on the 8-cell test suspension at a
fixed subset of targets, computed by the CPU ORACLE (oracle/rbc3d_oracle.c).

These are NOT reference-generated vectors: the Fortran + MPI + PETSc + FFTW reference cannot be built or run in this
image, so parity stays "unpinned" in the sense of the task contract.  The fixture pins the oracle
against accidental drift (tests/test_oracle_golden.py, CPU) and lets the GPU suite check the CUDA path against numbers
that do not depend on the oracle being rebuilt on the GPU box (tests/test_gpu_golden.py).

    python scripts/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle  # noqa: E402
from tests.util import C1_RHS, C2_MATVEC, small_suspension  # noqa: E402


def main():
    sus = small_suspension(2)                       # seed 161269, 8 cells, 36 x 72 points each
    orc = oracle.Oracle(sus.Lb).set_cells(sus)
    n = sus.npoint
    idx = np.arange(0, n, 97)[:256]                 # fixed subset of targets (every 97th point)
    out = {"idx": idx.astype(np.int64), "Lb": sus.Lb, "rc": np.array(orc.rc), "Nb": np.array(orc.Nb)}
    tl = orc.cell_targets()
    for name, c1, c2 in (("matvec", 0.0, C2_MATVEC), ("rhs", C1_RHS, 0.0)):
        v = orc.apply_cells(c1, c2, tl)
        out["v_" + name] = v[:, idx].copy()
        out["norm_" + name] = np.array(np.linalg.norm(v))
    out["cell_ids"] = orc.cell_ids(sus.x)[idx].astype(np.int32)
    path = os.path.join(ROOT, "tests", "golden", "oracle_8cells.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
