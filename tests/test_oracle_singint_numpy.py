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


# ---- Spline_FindProjection (ModSpline.F90:203-269) and the whole of AddIntOnRbcs (ModIntOnRbcs.F90:25-201) ----------
def polar_patch_map(th0, phi0, dth, dphi):
    s1 = np.array([np.cos(th0) * np.cos(phi0), np.cos(th0) * np.sin(phi0), -np.sin(th0)])
    s2 = np.array([-np.sin(phi0), np.cos(phi0), 0.0])
    x0 = np.array([np.sin(th0) * np.cos(phi0), np.sin(th0) * np.sin(phi0), np.cos(th0)])
    x = np.sin(dth) * np.cos(dphi) * s1 + np.sin(dth) * np.sin(dphi) * s2 + np.cos(dth) * x0
    x = x / np.sqrt(x @ x)
    th = np.arccos(max(-1.0, min(1.0, x[2])))
    phi = np.arctan2(x[1], x[0])
    return th, phi + 2 * PI if phi < 0 else phi


def find_projection(spx, xtar, th0, phi0):
    nth, nphi = 2, 8
    m, n = spx.shape[3], spx.shape[2]
    x0 = spline_interp(spx, th0, phi0)
    h = max(2 * PI / m, 2 * PI / n)
    for _ in range(3):
        thL = np.array([(i + 1) * h / nth for i in range(nth)])
        phL = np.arange(nphi) * 2 * PI / nphi
        thP, phP = polar_patch_build(th0, phi0, thL, phL)
        xy = [(0.0, 0.0)]
        d2 = [((spline_interp(spx, th0, phi0) - xtar) ** 2).sum()]
        for iphi in range(nphi):
            for ith in range(nth):
                xy.append((thL[ith] * np.cos(phL[iphi]), thL[ith] * np.sin(phL[iphi])))
                d2.append(((spline_interp(spx, thP[ith, iphi], phP[ith, iphi]) - xtar) ** 2).sum())
        xy, d2 = np.array(xy), np.array(d2)
        U = np.stack([np.ones(len(xy)), xy[:, 0], xy[:, 1], xy[:, 0] ** 2, xy[:, 0] * xy[:, 1], xy[:, 1] ** 2], axis=1)
        a0, a1, a2, a11, a12, a22 = np.linalg.solve(U.T @ U, U.T @ d2)          # QuadFit_2D: normal equations (dposv)
        det = 4 * a11 * a22 - a12 * a12                                            # Min_Quad_2D
        xm = np.array([(2 * a22 * (-a1) - a12 * (-a2)) / det, (-a12 * (-a1) + 2 * a11 * (-a2)) / det]) if det > 0 \
            else np.zeros(2)
        thMin, phiMin = polar_patch_map(th0, phi0, np.sqrt(xm @ xm), np.arctan2(xm[1], xm[0]))
        xMin = spline_interp(spx, thMin, phiMin)
        if ((xMin - xtar) ** 2).sum() > d2[0]:
            break
        th0, phi0, x0, h = thMin, phiMin, xMin, 0.5 * h
    return th0, phi0, x0


def add_int_on_rbcs_target(sus, tabs, c1, c2, xi, A_i, cell_i=None, ilat_i=None, ilon_i=None):
    """One target of AddIntOnRbcs without the linear term; cell_i None = raw target (indx = -1)."""
    npc = sus.nlat * sus.nlon
    radius = PI / np.sqrt(float(sus.nlat))
    fw = sus.weighted(sus.f) if c1 != 0 else None
    gw = sus.weighted(sus.g) if c2 != 0 else None
    d = sus.x - xi[:, None]
    d = d - np.rint(d / sus.Lb[:, None]) * sus.Lb[:, None]
    r = np.sqrt((d ** 2).sum(0))
    v = np.zeros(3)
    nbr = {}                                                            # NbrRbcList: other cell -> (dist, point)
    for j in np.nonzero(r <= tabs.rc)[0]:
        cj, pj = divmod(j, npc)
        ilon_j, ilat_j = divmod(pj, sus.nlat)
        xx, rr = d[:, j], r[j]
        if cell_i is not None and cj == cell_i:
            mask = mask_func(dist_on_sphere(sus.th[ilat_i], sus.phi[ilon_i], sus.th[ilat_j], sus.phi[ilon_j]) / radius)
        else:
            mask = 0.0
            if cj not in nbr or rr < nbr[cj][0]:
                nbr[cj] = (rr, j)
        if c1 != 0:
            EA, EB = tabs.sl(rr)
            v = v + (1.0 - mask) * c1 / A_i * (EA * xx * (xx @ fw[:, j]) + EB * fw[:, j])
        if c2 != 0:
            v = v + (1.0 - mask) * c2 * sus.Bcoef[cj] / A_i * (tabs.dl(rr) * xx * (xx @ gw[:, j]) * (xx @ sus.a3[:, j]))
    if cell_i is not None:
        v = v + sing_int(sus, tabs, c1, c2 * sus.Bcoef[cell_i], cell_i, ilat_i, ilon_i) / A_i
    for cj, (_, j) in nbr.items():
        pj = j % npc
        ilon0, ilat0 = divmod(pj, sus.nlat)
        x0 = sus.x[:, j]
        xx = xi - x0
        xt = x0 + (xx - np.rint(xx / sus.Lb) * sus.Lb)
        th0, phi0, x0p = find_projection(sus.spx[cj], xt, sus.th[ilat0], sus.phi[ilon0])
        v = v + nearsing_int(sus, tabs, c1, c2 * sus.Bcoef[cj], cj, xt, x0p, th0, phi0) / A_i
    return v


def linear_int(sus, c2):
    if c2 == 0:
        return np.zeros(3)
    npc = sus.nlat * sus.nlon
    vn = (sus.g * sus.a3).sum(0) * sus.dS()
    xv = (np.repeat(sus.Bcoef, npc)[None, :] * sus.x * vn[None, :]).sum(1)
    return c2 * (-8 * PI * np.prod(1.0 / sus.Lb) * xv)


def test_find_projection_equals_the_oracle(setup, oracle_lib):
    sus, orc, _ = setup
    rng = np.random.default_rng(2)
    for cell, ilat0, ilon0 in ((0, 5, 9), (4, 30, 50), (6, 0, 3)):
        p = cell * sus.nlat * sus.nlon + ilon0 * sus.nlat + ilat0
        xtar = sus.x[:, p] + 0.1 * sus.a3[:, p] + 0.03 * rng.normal(size=3)
        th, ph, x0 = find_projection(sus.spx[cell], xtar, sus.th[ilat0], sus.phi[ilon0])
        th_o, ph_o, x0_o = oracle_lib.Oracle.find_projection(sus.spx[cell], xtar, sus.th[ilat0], sus.phi[ilon0])
        assert abs(th - th_o) < 1e-9 and abs(ph - ph_o) < 1e-9 and np.abs(x0 - np.asarray(x0_o)).max() < 1e-9


@pytest.mark.parametrize("c1,c2", [(C1_RHS, 0.0), (0.0, C2_MATVEC), (C1_RHS, C1_RHS)])
def test_add_int_on_rbcs_equals_the_oracle(oracle_lib, c1, c2):
    """The whole of AddIntOnRbcs for a handful of targets of two nearly touching cells (pair sum with masks, singular,
    neighbour list + projection + near-singular incl. the jump branch, linear term) and for raw targets next to a
    surface -- second restatement vs the C oracle."""
    sus = util.close_pair_suspension(gap=0.004, seed=7)
    orc = oracle_lib.Oracle(sus.Lb).set_cells(sus)
    tabs = Tables(orc.alpha, orc.rc)
    npc = sus.nlat * sus.nlon
    # the points of cell 0 closest to cell 1, plus two ordinary ones
    d = sus.x[:, :npc, None] - sus.x[:, None, npc::7]
    close = np.argsort(np.sqrt((d ** 2).sum(0)).min(1))[:3]
    idx = np.concatenate([close, [100, npc + 777]])
    act = np.zeros(sus.npoint, np.int32)
    act[idx] = 1
    ref = orc.add_int_on_rbcs(c1, c2, orc.cell_targets(active=act))
    lin = linear_int(sus, c2)
    for i in idx:
        cell, p = divmod(i, npc)
        ilon, ilat = divmod(p, sus.nlat)
        A = sus.Acoef[cell]
        mine = add_int_on_rbcs_target(sus, tabs, c1, c2, sus.x[:, i], A, cell, ilat, ilon) + lin / A
        assert np.linalg.norm(mine - ref[:, i]) < 1e-9 * np.linalg.norm(ref[:, i])
    xr = sus.x[:, close[:2]] + 0.002 * sus.a3[:, close[:2]]             # raw targets just outside cell 0 (and near cell 1)
    ref = orc.add_int_on_rbcs(c1, c2, orc.make_targets(xr))
    for k in range(xr.shape[1]):
        mine = add_int_on_rbcs_target(sus, tabs, c1, c2, xr[:, k], 2.0) + lin / 2.0
        assert np.linalg.norm(mine - ref[:, k]) < 1e-9 * np.linalg.norm(ref[:, k])
