"""Tri_Int_Duffy (ModIntOnWalls.F90:373-465) and the direct loop of AddIntOnWalls (:80-126) restated a second time in
NumPy, straight from the Fortran, and compared with the C oracle (the regular 7-point rule and MinDistToTri already have
independent checks in tests/test_oracle_walls.py).  Uses the lerp tables of tests/test_oracle_singint_numpy.py."""
import numpy as np
import pytest

from rbc3d_b200 import synth
from tests.test_oracle_singint_numpy import Tables
from tests.util import C1_RHS

LB = np.array([10.5, 10.5, 8.0])


def tri_int_duffy(tabs, x, f, xtar, s0, t0):
    """x, f: (3 corners, 3).  -> rhs (3), lhs (3 corners, 3, 3)."""
    xg, wg = np.polynomial.legendre.leggauss(4)                          # GauLeg(0, 1, 4)
    rG, wG = 0.5 * (xg + 1), 0.5 * wg
    x0 = (1 - s0 - t0) * x[0] + s0 * x[1] + t0 * x[2]
    f0 = (1 - s0 - t0) * f[0] + s0 * f[1] + t0 * f[2]
    rhs, lhs = np.zeros(3), np.zeros((3, 3, 3))
    ref = {0: ((0.0, 0.0), (1.0, 0.0)), 1: ((1.0, 0.0), (0.0, 1.0)), 2: ((0.0, 1.0), (0.0, 0.0))}
    for n in range(3):
        x1, x2, f1, f2 = x[n], x[(n + 1) % 3], f[n], f[(n + 1) % 3]
        detJ = np.linalg.norm(np.cross(x1 - x0, x2 - x0))
        (s1, t1), (s2, t2) = ref[n]
        for i in range(4):
            for j in range(4):
                s = rG[i]
                t = s * rG[j]
                xq = (1 - s) * x0 + (s - t) * x1 + t * x2
                ds = wG[i] * wG[j] * detJ * s
                fq = ds * ((1 - s) * f0 + (s - t) * f1 + t * f2)
                xx = xtar - xq
                EA, EB = tabs.sl(np.sqrt(xx @ xx))
                rhs = rhs + (EA * xx * (xx @ fq) + EB * fq)
                K = ds * (EA * np.outer(xx, xx) + EB * np.eye(3))
                sg = (1 - s) * s0 + (s - t) * s1 + t * s2
                tg = (1 - s) * t0 + (s - t) * t1 + t * t2
                lhs[0] += (1 - sg - tg) * K
                lhs[1] += sg * K
                lhs[2] += tg * K
    return rhs, lhs


def tri_int_regular(tabs, x, f, xtar, rst, w):
    detJ = np.linalg.norm(np.cross(x[1] - x[0], x[2] - x[0]))            # 2 TriArea
    rhs = np.zeros(3)
    for (s, t), wq in zip(rst, w):
        xq = (1 - s - t) * x[0] + s * x[1] + t * x[2]
        fq = wq * detJ * ((1 - s - t) * f[0] + s * f[1] + t * f[2])
        xx = xtar - xq
        EA, EB = tabs.sl(np.sqrt(xx @ xx))
        rhs = rhs + (EA * xx * (xx @ fq) + EB * fq)
    return rhs


@pytest.fixture(scope="module")
def setup(oracle_lib):
    orc = oracle_lib.Oracle(LB)
    return orc, Tables(orc.alpha, orc.rc)


def test_duffy_equals_the_oracle(setup):
    orc, tabs = setup
    rng = np.random.default_rng(12)
    for _ in range(6):
        x = rng.uniform(3, 4, size=(3, 3))
        f = rng.normal(size=(3, 3))
        s0, t0 = rng.dirichlet([1, 1, 1])[:2]                            # closest point inside the triangle ...
        if _ % 2:
            s0, t0 = (0.0, rng.uniform()) if _ % 4 == 1 else (rng.uniform(), 0.0)      # ... or on an edge
        x0 = (1 - s0 - t0) * x[0] + s0 * x[1] + t0 * x[2]
        nrm = np.cross(x[1] - x[0], x[2] - x[0])
        xtar = x0 + 0.05 * nrm / np.linalg.norm(nrm)
        rhs, lhs = tri_int_duffy(tabs, x, f, xtar, s0, t0)
        r_o, l_o = orc.tri_int(x, f, xtar, s0, t0)
        assert np.linalg.norm(rhs - r_o) < 1e-12 * np.linalg.norm(r_o)
        assert np.abs(lhs - l_o).max() < 1e-12 * np.abs(l_o).max()
        assert np.allclose(np.einsum("lij,lj->i", lhs, f), rhs, rtol=1e-12, atol=1e-15)   # rhs = sum_l lhs(l) f(l), :316-317


def test_add_int_on_walls_direct_loop_equals_the_oracle(setup, oracle_lib):
    """raw targets near a tube wall: every element whose centroid is in the 27 list cells and whose exact distance is
    <= rc, Duffy when closer than sqrt(area), the 7-point rule otherwise, translated by the periodic image of vertex 1."""
    orc, tabs = setup
    W = synth.make_walls(LB, [dict(radius=4.4, ntheta=24, nz=10)], wobble=0.04)
    orc.set_walls(W, ncell=0)
    rst, wq = np.zeros((7, 2)), np.zeros(7)
    oracle_lib.lib().orc_gq_tri7(rst.ctypes.data_as(oracle_lib.c_dp), wq.ctypes.data_as(oracle_lib.c_dp))
    xt = np.array([[5.25 + 4.2, 5.25, 3.1], [5.25, 5.25 - 4.33, 7.95], [5.25 + 3.0, 5.25 + 3.0, 0.02]]).T   # last two near the period seam
    ref = orc.add_int_on_walls(C1_RHS, orc.make_targets(xt)) * 2.0       # raw targets: Acoef = 2
    e2v = W.e2v_global()
    xc = W.x[:, e2v].mean(axis=1)                                        # centroids (slist_wall%x)
    Nc = np.array(orc.Nc)
    for k in range(xt.shape[1]):
        xi = xt[:, k]
        ci = np.mod(np.floor(xi * Nc / LB).astype(int), Nc)
        cc = np.mod(np.floor(xc * (Nc / LB)[:, None]).astype(int), Nc[:, None])
        dcell = np.abs(cc - ci[:, None])
        near = (np.minimum(dcell, Nc[:, None] - dcell) <= 1).all(axis=0)  # the 27 list cells (periodic)
        v = np.zeros(3)
        ndf = 0
        for e in np.nonzero(near)[0]:
            x = W.x[:, e2v[:, e]].T.copy()                               # (corner, comp)
            f = W.f[:, e2v[:, e]].T
            sh = np.rint((xi - x[0]) / LB) * LB                          # ModIntOnWalls.F90:105: image of vertex 1
            x = x + sh[None, :]
            d, s0, t0, _ = oracle_lib.Oracle.min_dist_to_tri(xi, x)
            if d > orc.rc:
                continue
            if d < W.epsDist[e]:
                ndf += 1
                v = v + tri_int_duffy(tabs, x, f, xi, s0, t0)[0]
            else:
                v = v + tri_int_regular(tabs, x, f, xi, rst, wq)
        assert np.linalg.norm(C1_RHS * v - ref[:, k]) < 1e-11 * np.linalg.norm(ref[:, k])
        assert k != 0 or ndf > 0
