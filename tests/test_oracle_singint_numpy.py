"""RBC_SingInt (ModRbcSingInt.F90:29-90) restated a second time, in NumPy and straight from the Fortran -- polar patch
(RbcPolarPatch_Create / PolarPatch_Build, ModPolarPatch.F90:27-148), bicubic Hermite interpolation (Spline_Interp,
ModSpline.F90:150-191), lerp tables (MaskFunc, ModBasicMath.F90:351-379; EwaldCoeff_SL / _DL, ModEwaldFunc.F90:86-178) --
and compared with the C oracle per target.  The physical identities check the singular quadrature only to the method's
0.3 %; this pins the oracle's arithmetic for it to round-off against a second, separately written restatement (not
against the reference: PARITY UNPINNED stands)."""
import numpy as np
import pytest
from scipy.special import erfc

from tests import util
from tests.util import C1_RHS, C2_MATVEC

PI = np.pi
NTAB = 8192


def mask_func(x):
    s_ = np.arange(NTAB + 1) / NTAB
    with np.errstate(divide="ignore", over="ignore", invalid="ignore"):
        ex = np.exp(2 * np.exp(-1.0 / s_) / (s_ - 1))
    ftab = np.where(s_ < 0.01, 1.0, np.where(s_ > 0.99, 0.0, ex))
    s = abs(x) * NTAB
    i = int(np.floor(s))
    return 0.0 if i >= NTAB else ftab[i] * (i + 1 - s) + ftab[i + 1] * (s - i)


class Tables:
    def __init__(self, alpha, rc):
        self.alpha, self.rc = alpha, rc
        rt = np.sqrt(PI / alpha) * (np.arange(NTAB + 1) * rc / NTAB)
        self.sl1, self.sl2 = erfc(rt), 2 / np.sqrt(alpha) * np.exp(-rt ** 2)
        self.dl1 = -8 / np.sqrt(PI) * (np.exp(-rt ** 2) * (1.5 * rt + rt ** 3) + 0.75 * np.sqrt(PI) * erfc(rt))
        self.r_eps = 1e-3 * np.sqrt(alpha / PI)

    def _lerp(self, tab, r):
        s = NTAB * r / self.rc
        i = int(np.floor(s))
        return None if i >= NTAB else tab[i] * (i + 1 - s) + tab[i + 1] * (s - i)

    def sl(self, r):
        if r < self.r_eps or self._lerp(self.sl1, r) is None:
            return 0.0, 0.0
        c1, c2 = self._lerp(self.sl1, r), self._lerp(self.sl2, r)
        ir = 1.0 / r
        return c1 * ir ** 3 + c2 * ir ** 2, c1 * ir - c2

    def dl(self, r):
        if r < self.r_eps or self._lerp(self.dl1, r) is None:
            return 0.0
        return self._lerp(self.dl1, r) / r ** 5


def spline_interp(sp, x, y):
    """sp: (4 [u, u1, u2, u12], nvar, nlon, 2 nlat) of one cell, theta index fastest (u(i, j, l) = sp[0][l][j][i])."""
    _, nvar, n, m = sp.shape
    hx, hy = 2 * PI / m, 2 * PI / n
    i1, j1 = int(np.floor(x / hx)), int(np.floor(y / hy))
    s, t = x / hx - i1, y / hy - j1
    i1, j1 = i1 % m, j1 % n
    i2, j2 = (i1 + 1) % m, (j1 + 1) % n
    cx = np.array([1 + s * s * (-3 + 2 * s), s * s * (3 - 2 * s), hx * s * (1 + s * (-2 + s)), hx * s * s * (-1 + s)])
    cy = np.array([1 + t * t * (-3 + 2 * t), t * t * (3 - 2 * t), hy * t * (1 + t * (-2 + t)), hy * t * t * (-1 + t)])
    u, u1, u2, u12 = sp
    f = np.zeros(nvar)
    for l in range(nvar):
        U = np.array([[u[l, j1, i1], u[l, j2, i1], u2[l, j1, i1], u2[l, j2, i1]],
                      [u[l, j1, i2], u[l, j2, i2], u2[l, j1, i2], u2[l, j2, i2]],
                      [u1[l, j1, i1], u1[l, j2, i1], u12[l, j1, i1], u12[l, j2, i1]],
                      [u1[l, j1, i2], u1[l, j2, i2], u12[l, j1, i2], u12[l, j2, i2]]])
        f[l] = cx @ (U @ cy)
    return f


def polar_patch(nlat):
    radius = PI / np.sqrt(float(nlat))
    nrad = 2 * int(round(radius / (PI / nlat)))
    nazm = 2 * nrad
    xg, wg = np.polynomial.legendre.leggauss(nrad)                    # GauLeg(0, radius, nrad)
    thL = 0.5 * radius * (xg + 1)
    w = 0.5 * radius * wg
    w = np.array([w[k] * np.sin(thL[k]) * (2 * PI / nazm) * mask_func(thL[k] / radius) for k in range(nrad)])
    phiL = np.arange(nazm) * 2 * PI / nazm
    return radius, nrad, nazm, thL, phiL, w


def polar_patch_build(th0, phi0, thL, phiL):
    A = np.array([[np.cos(th0), 0, np.sin(th0)], [0, 1, 0], [-np.sin(th0), 0, np.cos(th0)]])
    x = np.stack([np.sin(thL)[:, None] * np.cos(phiL)[None, :], np.sin(thL)[:, None] * np.sin(phiL)[None, :],
                  np.cos(thL)[:, None] * np.ones_like(phiL)[None, :]])
    x = np.einsum("ab,bij->aij", A, x)
    thG = np.arccos(np.clip(x[2], -1.0, 1.0))
    phiG = np.arctan2(x[1], x[0]) + phi0
    return thG, phiG - np.floor(phiG / (2 * PI)) * 2 * PI


def sing_int(sus, tabs, c1, c2, cell, ilat0, ilon0):
    """dv of RBC_SingInt for target (cell, ilat0, ilon0) (0-based); c2 is c2Mod = c2 * Bcoef of the caller."""
    radius, nrad, nazm, thL, phiL, w = polar_patch(sus.nlat)
    thG, phiG = polar_patch_build(sus.th[ilat0], sus.phi[ilon0], thL, phiL)
    xi = spline_interp(sus.spx[cell], sus.th[ilat0], sus.phi[ilon0])
    dv = np.zeros(3)
    for irad in range(nrad):
        for iazm in range(nazm):
            th_j, phi_j = thG[irad, iazm], phiG[irad, iazm]
            xx = spline_interp(sus.spx[cell], th_j, phi_j) - xi
            rr = np.sqrt((xx * xx).sum())
            if rr >= tabs.rc:
                continue
            if c1 != 0:
                fj = w[irad] * spline_interp(sus.spF[cell], th_j, phi_j)
                EA, EB = tabs.sl(rr)
                dv = dv + c1 * (EA * xx * (xx @ fj) + EB * fj)
            if c2 != 0:
                gj = w[irad] * spline_interp(sus.spG[cell], th_j, phi_j)
                a3j = spline_interp(sus.spa3[cell], th_j, phi_j)
                dv = dv + c2 * (tabs.dl(rr) * xx * (xx @ gj) * (xx @ a3j))
    return dv


@pytest.fixture(scope="module")
def setup(oracle_lib):
    sus = util.small_suspension(2)                                      # 8 cells, 36 x 72 points each
    orc = oracle_lib.Oracle(sus.Lb).set_cells(sus)
    return sus, orc, Tables(orc.alpha, orc.rc)


def test_patch_tables_equal_the_oracle(setup):
    sus, orc, _ = setup
    radius, nrad, nazm, thL, phiL, w = polar_patch(sus.nlat)
    r_o, nrad_o, nazm_o, thG_o, phiG_o, w_o = orc.patch
    assert (nrad, nazm) == (nrad_o, nazm_o) == (12, 24) and abs(radius - r_o) < 1e-15
    assert np.allclose(w, np.asarray(w_o)[:nrad], rtol=1e-12, atol=1e-18)
    thG_o = np.asarray(thG_o).reshape(sus.nlon, sus.nlat, nazm, nrad)   # [nlon][nlat][nazm][nrad] (oracle header)
    phiG_o = np.asarray(phiG_o).reshape(sus.nlon, sus.nlat, nazm, nrad)
    for ilat0, ilon0 in ((0, 0), (17, 40), (35, 71)):
        thG, phiG = polar_patch_build(sus.th[ilat0], sus.phi[ilon0], thL, phiL)
        assert np.abs(thG.T - thG_o[ilon0, ilat0]).max() < 1e-13
        d = np.abs(phiG.T - phiG_o[ilon0, ilat0])
        assert np.minimum(d, 2 * PI - d).max() < 1e-12


@pytest.mark.parametrize("c1,c2", [(C1_RHS, 0.0), (0.0, C2_MATVEC), (C1_RHS, C1_RHS)])
def test_sing_int_equals_the_oracle(setup, c1, c2):
    sus, orc, tabs = setup
    for cell, ilat0, ilon0 in ((0, 0, 0), (3, 17, 40), (7, 35, 71), (5, 8, 13)):    # both poles' rows included
        c2m = c2 * sus.Bcoef[cell]                                      # c2Mod, ModIntOnRbcs.F90:116
        ref = orc.sing_int(c1, c2m, cell, ilat0 + 1, ilon0 + 1)
        mine = sing_int(sus, tabs, c1, c2m, cell, ilat0, ilon0)
        assert np.linalg.norm(mine - ref) < 1e-11 * np.linalg.norm(ref)


# ---- RBC_NearSingInt (ModRbcSingInt.F90:103-311) restated the same way ----------------------------------------------
def dist_on_sphere(th0, phi0, th1, phi1):
    d = np.cos(th0 - th1) - np.sin(th0) * np.sin(th1) * (1.0 - np.cos(phi0 - phi1))
    return np.arccos(min(1.0, max(-1.0, d)))


def find_points(th0, phi0, r0, ths, phis):
    """PolarPatch_FindPoints (ModPolarPatch.F90:160-207) -> list of 0-based (ilat, ilon)."""
    eps = 1e-10
    nphi = len(phis)
    ih = 1.0 / (2 * PI / nphi)
    out = []
    for i in range(len(ths)):
        if ths[i] <= th0 - r0:
            continue
        if ths[i] >= th0 + r0:
            break
        dphi = 1.0 - (np.cos(ths[i] - th0) - np.cos(r0)) / ((np.sin(ths[i]) + eps) * (np.sin(th0) + eps))
        dphi = np.arccos(max(-1.0, min(1.0, dphi)))
        if dphi > PI - eps:
            jmin, jmax = 1, nphi
        else:
            jmin = int(np.ceil((phi0 - dphi - phis[0]) * ih)) + 1
            jmax = int(np.floor((phi0 + dphi - phis[0]) * ih)) + 1
        out += [(i, (j - 1) % nphi) for j in range(jmin, jmax + 1)]
    return out


def gauleg_sinh(xmin, xmax, a, b, n):
    xm, xl = 0.5 * (xmin + xmax), 0.5 * (xmax - xmin)
    a0, b0 = (a - xm) / xl, b * xl
    u1, u2 = np.arcsinh((1 + a0) / b0), np.arcsinh((1 - a0) / b0)
    mu, eta = 0.5 * (u1 + u2), 0.5 * (u1 - u2)
    s, w = np.polynomial.legendre.leggauss(n)
    x = a0 + b0 * np.sinh(mu * s - eta)
    w = w * b0 * mu * np.cosh(mu * s - eta)
    return xm + xl * x, xl * w


def nearsing_subtract(sus, tabs, c1, c2, cell, xi, th0, phi0, radPat):
    npc = sus.nlat * sus.nlon
    fw = sus.weighted(sus.f) if c1 != 0 else None
    gw = sus.weighted(sus.g) if c2 != 0 else None
    dv = np.zeros(3)
    for ilat, ilon in find_points(th0, phi0, radPat, sus.th, sus.phi):
        p = cell * npc + ilon * sus.nlat + ilat
        xx = sus.x[:, p] - xi
        rr = np.sqrt((xx ** 2).sum())
        if rr > tabs.rc:
            continue
        mask = mask_func(dist_on_sphere(th0, phi0, sus.th[ilat], sus.phi[ilon]) / radPat)
        if c1 != 0:
            EA, EB = tabs.sl(rr)
            dv = dv - c1 * mask * (EA * xx * (xx @ fw[:, p]) + EB * fw[:, p])
        if c2 != 0:
            dv = dv - c2 * mask * (tabs.dl(rr) * xx * (xx @ gw[:, p]) * (xx @ sus.a3[:, p]))
    return dv


def nearsing_readd(sus, tabs, c1, c2, cell, xi, x0, th0, phi0, radPat):
    dist = np.sqrt(((xi - x0) ** 2).sum())
    sizePat = radPat * np.sqrt(sus.area[cell] / (4 * PI))
    nrad = 16
    nazm = 2 * nrad
    if dist > np.finfo(float).tiny:
        thP, wP = gauleg_sinh(0.0, radPat, 0.0, dist * (radPat / sizePat), nrad)
    else:
        xg, wg = np.polynomial.legendre.leggauss(nrad)
        thP, wP = 0.5 * radPat * (xg + 1), 0.5 * radPat * wg
    wP = np.array([wP[k] * np.sin(thP[k]) * (2 * PI / nazm) * mask_func(thP[k] / radPat) for k in range(nrad)])
    thG, phiG = polar_patch_build(th0, phi0, thP, np.arange(nazm) * 2 * PI / nazm)
    dv = np.zeros(3)
    for irad in range(nrad):
        for iazm in range(nazm):
            th_j, phi_j = thG[irad, iazm], phiG[irad, iazm]
            xx = spline_interp(sus.spx[cell], th_j, phi_j) - xi
            rr = np.sqrt((xx ** 2).sum())
            if rr > tabs.rc:
                continue
            if c1 != 0:
                fj = wP[irad] * spline_interp(sus.spF[cell], th_j, phi_j)
                EA, EB = tabs.sl(rr)
                dv = dv + c1 * (EA * xx * (xx @ fj) + EB * fj)
            if c2 != 0:
                a3j = spline_interp(sus.spa3[cell], th_j, phi_j)
                gj = wP[irad] * spline_interp(sus.spG[cell], th_j, phi_j)
                dv = dv + c2 * (tabs.dl(rr) * xx * (xx @ gj) * (xx @ a3j))
    return dv


def nearsing_int(sus, tabs, c1, c2, cell, xi, x0, th0, phi0):
    radPat = PI / np.sqrt(float(sus.nlat))
    a30 = spline_interp(sus.spa3[cell], th0, phi0)
    dist = a30 @ (xi - x0)
    if dist > 2 * sus.meshSize[cell]:
        return np.zeros(3)
    sizePat = radPat * np.sqrt(sus.area[cell] / (4 * PI))
    dist1 = np.copysign(0.01 * sizePat, dist)
    dv = nearsing_subtract(sus, tabs, c1, c2, cell, xi, th0, phi0, radPat)
    if abs(dist) >= abs(dist1):
        return dv + nearsing_readd(sus, tabs, c1, c2, cell, xi, x0, th0, phi0, radPat)
    dv1 = nearsing_readd(sus, tabs, c1, c2, cell, x0 + dist1 * a30, x0, th0, phi0, radPat)
    dv0 = nearsing_readd(sus, tabs, c1, c2, cell, x0.copy(), x0, th0, phi0, radPat)
    if c2 != 0:
        g0 = spline_interp(sus.spG[cell], th0, phi0) / spline_interp(sus.spdetj[cell], th0, phi0)[0]
        dv0 = dv0 + (c2 * 4 * PI * g0 if dist > 0 else -c2 * 4 * PI * g0)
    return dv + dv0 + dist / dist1 * (dv1 - dv0)


@pytest.mark.parametrize("c1,c2", [(C1_RHS, 0.0), (0.0, C2_MATVEC), (C1_RHS, C1_RHS)])
def test_nearsing_int_equals_the_oracle(setup, oracle_lib, c1, c2):
    """every branch: far (> 2 meshSize: zero), regular sinh rule on either side of the surface, and the jump-interpolation
    branch |dist| < 0.01 sizePat on either side (+- 4 pi c2 g0); projection foot off the mesh, near a pole too."""
    sus, orc, tabs = setup
    for cell, th0, phi0 in ((2, 1.234, 4.321), (6, 0.09, 0.5), (1, 3.0, 6.1)):
        x0 = spline_interp(sus.spx[cell], th0, phi0)
        a30 = spline_interp(sus.spa3[cell], th0, phi0)
        c2m = c2 * sus.Bcoef[cell]
        for d in (0.05, 0.003, -0.003, -0.05, 0.2, 0.4):
            xi = x0 + d * a30
            ref = orc.nearsing_int(c1, c2m, cell, xi, x0, th0, phi0)
            mine = nearsing_int(sus, tabs, c1, c2m, cell, xi, x0, th0, phi0)
            if d > 2 * sus.meshSize[cell]:
                assert not ref.any() and not mine.any()
            else:
                assert np.linalg.norm(mine - ref) < 1e-10 * np.linalg.norm(ref)
