"""CUDA path against the committed oracle-generated fixture (no oracle call at run time)."""
import os

import numpy as np
import pytest

from tests.util import C1_RHS, C2_MATVEC, rel_l2, small_suspension

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_8cells.npz")


def test_gpu_operator_against_committed_vectors():
    from rbc3d_b200.ewald import EwaldOperator
    g = np.load(GOLD)
    sus = small_suspension(2)
    op = EwaldOperator(sus.Lb)
    op.set_suspension(sus)
    idx = g["idx"]
    _, cid, _, _ = op.cell_list()
    assert np.array_equal(cid[idx], g["cell_ids"])                           # bit-exact
    for name, c1, c2 in (("matvec", 0.0, C2_MATVEC), ("rhs", C1_RHS, 0.0)):
        v = op.apply(c1, c2)
        assert rel_l2(v[:, idx], g["v_" + name]) < 1e-10                     # north-star tolerance
        assert abs(np.linalg.norm(v) / float(g["norm_" + name]) - 1.0) < 1e-10
    op.close()
