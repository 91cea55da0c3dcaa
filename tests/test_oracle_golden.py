"""The oracle against the committed fixture tests/golden/oracle_8cells.npz (made by scripts/make_golden.py FROM THE
ORACLE: it guards against drift of the restatement, it does not pin it to the reference -- see the fixture script)."""
import os

import numpy as np

from tests.util import C1_RHS, C2_MATVEC, rel_l2, small_suspension

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_8cells.npz")


def test_oracle_reproduces_the_committed_vectors():
    from oracle import oracle
    g = np.load(GOLD)
    sus = small_suspension(2)
    orc = oracle.Oracle(sus.Lb).set_cells(sus)
    assert list(g["Nb"]) == orc.Nb and abs(float(g["rc"]) - orc.rc) < 1e-15
    idx = g["idx"]
    assert np.array_equal(orc.cell_ids(sus.x)[idx], g["cell_ids"])          # integer work: bit-exact
    tl = orc.cell_targets()
    for name, c1, c2 in (("matvec", 0.0, C2_MATVEC), ("rhs", C1_RHS, 0.0)):
        v = orc.apply_cells(c1, c2, tl)
        # thread count changes the OpenMP summation order of the mesh: equal up to rounding
        assert rel_l2(v[:, idx], g["v_" + name]) < 1e-12
        assert abs(np.linalg.norm(v) / float(g["norm_" + name]) - 1.0) < 1e-12
