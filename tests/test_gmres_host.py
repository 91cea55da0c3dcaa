"""CPU checks of the host-side solver harness (rbc3d_b200/gmres.py): SH packing of Glob_Sph_Trans and the restated
PETSc-default GMRES.  The operator itself is not called here (no GPU)."""
import numpy as np

from rbc3d_b200 import gmres as G
from rbc3d_b200 import sphere


def test_sh_analysis_synthesis_roundtrip():
    nlat0, nlat, nlon = 6, 18, 36
    sh = G.ShTransform(nlat, nlon, nlat0)
    rng = np.random.default_rng(0)
    a = np.zeros((2, nlat0, nlat0))
    b = np.zeros_like(a)
    for m in range(nlat0):
        a[:, m, m:] = rng.uniform(-1, 1, (2, nlat0 - m))
        if m:
            b[:, m, m:] = rng.uniform(-1, 1, (2, nlat0 - m))
    f = sh.synth(a, b)
    a2, b2 = sh.anal(f)
    assert np.abs(a2 - a).max() < 1e-12 and np.abs(b2 - b).max() < 1e-12
    # shsgs convention: the m = 0 coefficients enter with a factor 1/2, the others with (cos, -sin)
    th, phi, _ = sphere.gauss_grid(nlat, nlon)
    pb = sphere._pbar(nlat0, np.cos(th))
    a1 = np.zeros((nlat0, nlat0)); b1 = np.zeros_like(a1)
    a1[0, 2] = 1.0
    assert np.allclose(sh.synth(a1, b1), 0.5 * pb[0, 2][None, :] * np.ones((nlon, 1)), atol=1e-13)
    a1[:] = 0; b1[3, 4] = 1.0
    assert np.allclose(sh.synth(a1, b1), -np.sin(3 * phi)[:, None] * pb[3, 4][None, :], atol=1e-13)


def test_glob_sph_trans_packing_and_projection():
    ncell, nlat0, nlat, nlon = 3, 4, 12, 24
    T = G.GlobSphTrans(ncell, nlat, nlon, nlat0)
    assert T.dof == ncell * 3 * nlat0 ** 2                     # ModVelSolver.F90:66
    rng = np.random.default_rng(1)
    c = rng.uniform(-1, 1, T.dof)
    v = T.four_to_phys(c)
    assert v.shape == (3, ncell * nlat * nlon)
    assert np.abs(T.phys_to_four(v) - c).max() < 1e-12
    # first entries of a cell: a(m=0, n=0) of the three components (components interleaved, ModVelSolver.F90:676-688)
    c0 = np.zeros(T.dof); c0[1] = 2.0
    v0 = T.four_to_phys(c0)
    npc = nlat * nlon
    assert np.allclose(v0[1, :npc], 2.0 * 0.5 * np.sqrt(0.5)) and np.allclose(v0[0], 0) and np.allclose(v0[1, npc:], 0)
    # analysis of a non-band-limited field is a projection
    w = rng.uniform(-1, 1, (3, ncell * npc))
    p1 = T.four_to_phys(T.phys_to_four(w))
    assert np.abs(T.four_to_phys(T.phys_to_four(p1)) - p1).max() < 1e-12


def test_gmres_matches_dense_solve_and_restarts():
    rng = np.random.default_rng(2)
    n = 120
    A = np.eye(n) + 0.3 * rng.standard_normal((n, n)) / np.sqrt(n)
    b = rng.standard_normal(n)
    x, it, hist = G.gmres(lambda v: A @ v, b, x0=np.zeros(n), rtol=1e-11)
    assert np.linalg.norm(A @ x - b) < 2e-11 * np.linalg.norm(b)
    assert it == len(hist) - 1 and it < 40
    assert all(hist[i + 1] <= hist[i] * (1 + 1e-12) for i in range(len(hist) - 1))   # minimal residual
    # recurrence residual = true residual (no preconditioner) while no restart happened
    x5, it5, h5 = G.gmres(lambda v: A @ v, b, x0=np.zeros(n), rtol=1e-30, maxit=5)
    assert it5 == 5 and abs(np.linalg.norm(b - A @ x5) - h5[-1]) < 1e-12
    # harder system: needs restarts (restart 7), still converges, iteration count monotone in rtol
    B = np.eye(n) + 0.9 * rng.standard_normal((n, n)) / np.sqrt(n)
    xr, itr, hr = G.gmres(lambda v: B @ v, b, x0=np.zeros(n), rtol=1e-9, restart=7, maxit=2000)
    assert np.linalg.norm(B @ xr - b) < 1e-8 * np.linalg.norm(b) and itr > 7
    # nonzero initial guess: starting from the solution takes zero iterations
    xs, its, hs = G.gmres(lambda v: A @ v, b, x0=x, rtol=1e-10)
    assert its == 0


def test_gmres_against_scipy_iteration_history():
    import scipy.sparse.linalg as spla
    rng = np.random.default_rng(3)
    n = 80
    A = np.eye(n) + 0.4 * rng.standard_normal((n, n)) / np.sqrt(n)
    b = rng.standard_normal(n)
    res = []
    spla.gmres(A, b, x0=np.zeros(n), rtol=1e-10, restart=30, maxiter=10, callback=lambda r: res.append(r),
               callback_type="pr_norm")
    _, it, hist = G.gmres(lambda v: A @ v, b, x0=np.zeros(n), rtol=1e-10)
    k = min(len(res), it)
    assert k >= 5
    # scipy reports the relative preconditioned residual after every inner iteration
    assert np.allclose(np.array(hist[1:k + 1]) / np.linalg.norm(b), res[:k], rtol=1e-6)


def test_gmres_restatement_against_scipy():
    """Independent implementation: scipy.sparse.linalg.gmres (restart 30, no preconditioner) must produce the same
    residual history and solution as the KSPGMRES restatement -- in exact arithmetic the history does not depend on the
    orthogonalisation, so classical Gram-Schmidt (PETSc's default, ours) and SciPy's agree to round-off on a
    well-conditioned second-kind operator; the restart boundary (30) is crossed."""
    from scipy.sparse.linalg import LinearOperator, gmres as sp_gmres
    from rbc3d_b200.gmres import gmres
    rng = np.random.default_rng(5)
    n = 300
    K = rng.normal(size=(n, n)) / np.sqrt(n)
    A = np.eye(n) + 0.9 * K                       # spectrum in a disc of radius ~0.9 around 1: slow enough to restart
    b = rng.normal(size=n)
    x, it, hist = gmres(lambda u: A @ u, b, rtol=1e-10, restart=30, maxit=200)
    assert it > 30                                # crossed a restart
    res = []
    xs, info = sp_gmres(LinearOperator((n, n), matvec=lambda u: A @ u), b, rtol=1e-10, atol=0.0, restart=30,
                        maxiter=20, callback=lambda r: res.append(r), callback_type="pr_norm")
    assert info == 0
    res = np.array(res) * np.linalg.norm(b)       # SciPy reports ||r|| / ||b||
    m = min(len(res), len(hist) - 1)
    assert m >= it - 1
    assert np.allclose(res[:m], hist[1:m + 1], rtol=1e-6)
    assert np.linalg.norm(x - xs) < 1e-8 * np.linalg.norm(x)
