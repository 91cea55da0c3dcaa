"""One draw of the cell placement of examples/carotid_web (rbc3d_b200.cases.carotid_place_cells: the rejection sampling
of carotid_initcond.F90 restated; NumPy PCG64 seed 112) -> rbc3d_b200/data/carotid_web_cells.npz (72 centres and rotation
matrices, 7 KB).  The sampling takes minutes at 72 cells (as the init program's does), the GPU tests must not.

    python scripts/make_golden_carotid_cells.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    from rbc3d_b200 import cases
    W, Lb = cases.carotid_web_walls()
    t0 = time.time()
    centres, rots = cases.carotid_place_cells(W, 72, seed=112, progress=True)
    print("placed 72 cells in %.0f s" % (time.time() - t0))
    np.savez(os.path.join(ROOT, "rbc3d_b200", "data", "carotid_web_cells.npz"), centres=centres, rotations=rots,
             seed=112, Lb=Lb)
