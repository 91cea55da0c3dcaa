"""Shared helpers of the test-suite (suspension builders, error norms)."""
import numpy as np

from rbc3d_b200 import synth

C2_MATVEC = -1.0 / (4.0 * np.pi)   # cell GMRES matvec, ModVelSolver.F90:568-569
C1_RHS = 1.0 / (4.0 * np.pi)       # Compute_Rhs, ModVelSolver.F90:465


def rel_l2(a, b):
    """relative L2 error over all targets and components (the north-star parity metric)."""
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def small_suspension(n_side=2, seed=161269, **kw):
    return synth.make_suspension(n_side, seed=seed, **kw)


def min_gap(sus, a=0, b=1):
    """smallest distance between mesh points of cells a and b (no periodic images)."""
    npc = sus.nlat * sus.nlon
    xa = sus.x[:, a * npc:(a + 1) * npc]
    xb = sus.x[:, b * npc:(b + 1) * npc]
    d2 = ((xa[:, :, None] - xb[:, None, :]) ** 2).sum(0)
    return float(np.sqrt(d2.min()))


def close_pair_suspension(gap=0.08, L=9.0, seed=7, extra=0):
    """Two (or 2+extra) cells in a box with the first two nearly touching (exercises the near-singular path,
    including the |dist| < 0.01*sizePat interpolation branch for very small gaps)."""
    c0 = np.array([L / 2 - 0.9, L / 2, L / 2])
    direction = np.array([1.0, 0.2, -0.1])
    direction /= np.linalg.norm(direction)
    sep = 3.0
    rng = np.random.default_rng(seed)
    others = rng.uniform(1.5, L - 1.5, size=(extra, 3))
    sus = None
    for _ in range(12):
        centers = np.vstack([c0, c0 + sep * direction] + [o for o in others])
        sus = synth.make_suspension(1, L=L, centers=centers, seed=seed)
        g = min_gap(sus)
        if abs(g - gap) < 1e-3:
            break
        sep -= (g - gap)
    return sus
