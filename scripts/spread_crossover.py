"""Spreading-kernel crossover: time PME_Distrib_Source with the pencil walk and with the source-block kernel for
suspensions of n_side^3 cells at the packing of the 4096-cell benchmark (run on a GPU box).  The result places
RBC3D_SPREAD_WALK_MIN (rbc3d_internal.h: Pme::swalk_min)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rbc3d_b200 import synth  # noqa: E402
from rbc3d_b200.ewald import EwaldOperator  # noqa: E402

C1 = 1.0 / (4 * np.pi)
out = []
for n_side in (3, 4, 5, 6, 8):
    sus = synth.make_suspension(n_side)
    row = {"cells": n_side ** 3, "points": int(sus.x.size // 3)}
    for mode, name in (("0", "walk"), ("1", "blocks")):
        os.environ["RBC3D_SPREAD_BLOCKS"] = mode
        op = EwaldOperator(sus.Lb)
        op.set_suspension(sus)
        best = {}
        for c1, c2, tag in ((C1, 0.0, "sl"), (0.0, -C1, "dl")):
            ts = []
            for _ in range(6):
                op.PME_Distrib_Source(c1, c2, cells=True)
                ts.append(op.timings()["spread"])
            best[tag] = min(ts[1:])
        row[name] = best
        row["Nb"] = list(op.Nb)
        op.close()
    out.append(row)
    print(json.dumps(row), flush=True)
