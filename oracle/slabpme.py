"""TEST INFRASTRUCTURE / DESIGN MODEL, NOT PRODUCT CODE (lives under oracle/ with the other CPU restatements).

NumPy model of the PME part of the operator with the SLAB-DECOMPOSED transform of the reference:
the executable specification of the multi-GPU PME transpose path (next round: cuFFT 2-D per z-slab + all-to-all over
NVLink + 1-D FFT per y-slab), checked on CPU against the oracle and, over two gloo ranks, against the single-rank
result (tests/test_slab_pme.py).  The CUDA library of this round sums full meshes with one all-reduce and transforms
redundantly on every rank; nothing under rbc3d_b200/ imports this module.

Restated here, each citing the reference:
* ``bspline_func``        BSplineFunc, ModBasicMath.F90:392-417 (vectorised over points)
* ``PmeModel.spread``     Distrib_Source, ModPME.F90:405-443 (weights w_x w_y w_z, periodic wrap, c1 f / c2 g (x) a3 B)
* ``PmeModel.modulus``    the B-spline factor bb of PME_Init, ModPME.F90:310-336
* ``PmeModel.multiply``   the Stokeslet / stresslet multipliers of PME_Transform, ModPME.F90:165-210, with the wave
                          vector q = (i/L1, +-j/L2, -+k/L3) of :287-303 (the z sign is flipped)
* transforms              ModPFFTW.F90:94-185: forward = r2c over (x, y) with e^{-i}, then z with e^{+i}; backward = z with
                          e^{-i}, y with e^{+i}, c2r over x (which drops the imaginary part of the x = 0 and x = Nx/2
                          bins AFTER the z and y transforms -- so that step is local to a z-plane and needs no partner
                          rank in the slab layout); both unnormalised
* ``slab_chunks``         Init_PFFTW, ModPFFTW.F90:56-89: chunk = N / R, ranks 1..mod(N, R) get one more (not rank 0)
* ``SlabRank``            Get_That / Get_R, ModPFFTW.F90:148-185: z-slabs <-> y-slabs by an all-to-all of
                          [component][z of the sender][y of the receiver][Nx/2+1] blocks
* ``PmeModel.interp``     Interp_Vel, ModPME.F90:450-489
"""
from __future__ import annotations

import numpy as np

PI = np.pi


def bspline_func(xc: np.ndarray, P: int):
    """-> imin (n,) int64, w (n, P): the P non-zero B-spline values at mesh points imin .. imin + P - 1."""
    xc = np.asarray(xc, dtype=float)
    imin = np.floor(xc).astype(np.int64) - (P - 1)
    u = (imin - (xc - P))[:, None] + np.arange(P)[None, :]          # u(1) = imin - (xc - P), u(j) = u(j-1) + 1
    w = np.zeros((xc.size, P))
    w[:, 0] = 1.0
    for pp in range(2, P + 1):
        for j in range(pp, 1, -1):                                   # w(j) uses the old w(j) and w(j-1)
            w[:, j - 1] = u[:, j - 1] / (pp - 1.0) * w[:, j - 1] + (pp - u[:, j - 1]) / (pp - 1.0) * w[:, j - 2]
        w[:, 0] = u[:, 0] / (pp - 1.0) * w[:, 0]
    return imin, w


def slab_chunks(N: int, R: int):
    """[(lo, hi)) of every rank for N planes over R ranks with the reference's remainder rule."""
    chunk = [N // R] * R
    for i in range(1, N % R + 1):                                    # chunk(i) = chunk(i) + 1 for i = 1 .. mod(N, R)
        chunk[i % R] += 1                                            # (i < R always: mod(N, R) <= R - 1)
    lo = np.concatenate([[0], np.cumsum(chunk)])
    return [(int(lo[r]), int(lo[r + 1])) for r in range(R)]


class PmeModel:
    """Single-rank PME in NumPy.  Meshes are [component][Nz][Ny][Nx] real and [component][Nz][Ny][Nx/2+1] complex
    (the layout of pme.cu); DL tensor components in the order tt(ii, jj) -> index ii + 3 jj (the oracle's)."""

    def __init__(self, Lb, Nb, alpha: float = 0.44, P: int = 8):
        self.Lb, self.Nb, self.alpha, self.P = np.asarray(Lb, dtype=float), [int(n) for n in Nb], float(alpha), int(P)
        Nx, Ny, Nz = self.Nb
        self.Nxh = Nx // 2 + 1
        self.vol = float(np.prod(self.Lb))
        bx, by, bz = (self.modulus(n, c) for n, c in ((Nx, self.Nxh), (Ny, Ny), (Nz, Nz)))
        self.bb = ((1.0 * bx[None, None, :]) * by[None, :, None]) * bz[:, None, None]      # [Nz][Ny][Nxh]
        self.bb[0, 0, 0] = 0.0

    def modulus(self, N: int, count: int) -> np.ndarray:
        P = self.P
        _, MP = bspline_func(np.array([P + np.finfo(float).eps]), P)                        # BSplineFunc(P + EPSILON)
        k = np.arange(count)[:, None]
        b = (MP[0, :P - 1][None, :] * np.exp(2j * PI * k * np.arange(P - 1)[None, :] / N)).sum(1)
        b = np.exp(2j * PI * np.arange(count) * (P - 1.0) / N) / b
        return np.abs(b) ** 2

    # -- Distrib_Source ---------------------------------------------------------------------------------------------
    def spread(self, x, c1=0.0, c2=0.0, f=None, g=None, a3=None, Bcoef=None, zrange=None):
        """-> (ff (3, Nz, Ny, Nx) or None, tt (9, Nz, Ny, Nx) or None).  zrange = (lo, hi): keep only those planes (the
        ``cycle`` of ModPME.F90:428-429) -- the arrays stay full size."""
        Nx, Ny, Nz = self.Nb
        P = self.P
        ih = np.array(self.Nb) / self.Lb
        im, wx = bspline_func(x[0] * ih[0], P)
        jm, wy = bspline_func(x[1] * ih[1], P)
        km, wz = bspline_func(x[2] * ih[2], P)
        ar = np.arange(P)
        i = np.mod(im[:, None] + ar, Nx)
        j = np.mod(jm[:, None] + ar, Ny)
        k = np.mod(km[:, None] + ar, Nz)
        w = wx[:, None, None, :] * wy[:, None, :, None] * wz[:, :, None, None]               # (n, k0, j0, i0)
        if zrange is not None:
            w = w * ((k >= zrange[0]) & (k < zrange[1]))[:, :, None, None]
        flat = ((k[:, :, None, None] * Ny + j[:, None, :, None]) * Nx + i[:, None, None, :]).reshape(-1)
        ff = tt = None
        if abs(c1) > 1e-10:
            ff = np.zeros((3, Nz * Ny * Nx))
            for d in range(3):
                np.add.at(ff[d], flat, ((c1 * w) * f[d][:, None, None, None]).reshape(-1))
            ff = ff.reshape(3, Nz, Ny, Nx)
        if abs(c2) > 1e-10:
            tt = np.zeros((9, Nz * Ny * Nx))
            for ii in range(3):
                for jj in range(3):
                    t = g[ii] * (a3[jj] * Bcoef)                                             # t = g (x) a3 Bcoef
                    np.add.at(tt[ii + 3 * jj], flat, ((c2 * w) * t[:, None, None, None]).reshape(-1))
            tt = tt.reshape(9, Nz, Ny, Nx)
        return ff, tt

    # -- transforms (single rank) ---------------------------------------------------------------------------------
    def forward(self, a):
        return np.fft.ifft(np.fft.rfft2(a, axes=(-2, -1)), axis=-3) * self.Nb[2]

    def backward(self, aC):
        Nx, Ny, _ = self.Nb
        return np.fft.irfft(np.fft.ifft(np.fft.fft(aC, axis=-3), axis=-2) * Ny, n=Nx, axis=-1) * Nx

    # -- PME_Transform multipliers on a range of y modes ----------------------------------------------------------------
    def multiply(self, ffC, ttC, jrange=None):
        """ffC (3, Nz, ny, Nxh) / ttC (9, Nz, ny, Nxh) for the y modes jrange = (lo, hi) -> vvC (3, Nz, ny, Nxh)."""
        Nx, Ny, Nz = self.Nb
        lo, hi = jrange if jrange is not None else (0, Ny)
        iL = 1.0 / self.Lb
        i = np.arange(self.Nxh)
        j = np.arange(lo, hi)
        k = np.arange(Nz)
        q = np.zeros((3, Nz, hi - lo, self.Nxh))
        q[0] = (i * iL[0])[None, None, :]
        q[1] = (np.where(j < Ny // 2, j, j - Ny) * iL[1])[None, :, None]
        q[2] = (-np.where(k < Nz // 2, k, k - Nz) * iL[2])[:, None, None]
        a = self.alpha
        q2t = PI * a * (q ** 2).sum(0)
        zero = q2t == 0.0
        q2t = np.where(zero, 1.0, q2t)
        e = np.exp(-q2t)
        phi0 = e / q2t
        phi1 = (e + phi0) / q2t
        vC = np.zeros((3, Nz, hi - lo, self.Nxh), dtype=complex)
        if ffC is not None:
            qt = np.sqrt(PI * a) * q
            dotp = (qt * ffC).sum(0)
            vC += 2 * a / self.vol * phi1 * (q2t * ffC - qt * dotp)
        if ttC is not None:
            T = ttC.reshape(3, 3, Nz, hi - lo, self.Nxh).transpose(1, 0, 2, 3, 4)            # T[ii][jj] = ttC[ii + 3 jj]
            tr = T[0, 0] + T[1, 1] + T[2, 2]
            qT = np.einsum("a...,ab...->b...", q, T)
            Tq = np.einsum("ab...,b...->a...", T, q)
            qTq = (q * Tq).sum(0)
            t = 1j * 4 * PI * a / self.vol * phi0 * (q * tr + qT + Tq)
            t = t - 1j * 8 * PI * PI * a * a / self.vol * phi1 * qTq * q
            vC -= t
        vC[:, zero] = 0.0
        return vC * self.bb[None, :, lo:hi, :]

    def transform(self, ff, tt):
        return self.backward(self.multiply(self.forward(ff) if ff is not None else None,
                                           self.forward(tt) if tt is not None else None))

    # -- Interp_Vel ---------------------------------------------------------------------------------------------------
    def interp(self, x, vv):
        Nx, Ny, Nz = self.Nb
        P = self.P
        ih = np.array(self.Nb) / self.Lb
        im, wx = bspline_func(x[0] * ih[0], P)
        jm, wy = bspline_func(x[1] * ih[1], P)
        km, wz = bspline_func(x[2] * ih[2], P)
        ar = np.arange(P)
        i, j, k = np.mod(im[:, None] + ar, Nx), np.mod(jm[:, None] + ar, Ny), np.mod(km[:, None] + ar, Nz)
        w = wx[:, None, None, :] * wy[:, None, :, None] * wz[:, :, None, None]
        flat = (k[:, :, None, None] * Ny + j[:, None, :, None]) * Nx + i[:, None, None, :]
        return np.stack([(w * vv[d].reshape(-1)[flat]).sum(axis=(1, 2, 3)) for d in range(3)])


    def interp_slab(self, x, vslab, zlo: int, zhi: int):
        """Interp_Vel on a rank that holds only the planes zlo - P .. zhi - 1 of the velocity mesh (its z-slab plus the P
        planes received from the -z neighbour, Update_Buff_Vel, ModPME.F90:354-396): vslab (3, P + zhi - zlo, Ny, Nx),
        plane 0 = mesh plane zlo - P (periodic).  Valid for the targets the rank owns, floor(z Nb3 / Lb3) mod Nb3 in
        [zlo, zhi) (SetActiveFlag); the index arithmetic is the one of ModPME.F90:470-473."""
        Nx, Ny, Nz = self.Nb
        P = self.P
        ih = np.array(self.Nb) / self.Lb
        im, wx = bspline_func(x[0] * ih[0], P)
        jm, wy = bspline_func(x[1] * ih[1], P)
        km, wz = bspline_func(x[2] * ih[2], P)
        ar = np.arange(P)
        i, j = np.mod(im[:, None] + ar, Nx), np.mod(jm[:, None] + ar, Ny)
        lb = zlo - P                                                    # lbound(vv, 3)
        k = lb + np.mod(km[:, None] + ar - lb, Nz)
        ok = k <= zhi - 1                                               # "if (k > ubound(vv,3)) cycle"
        kk = np.where(ok, k - lb, 0)
        w = wx[:, None, None, :] * wy[:, None, :, None] * (wz * ok)[:, :, None, None]
        flat = (kk[:, :, None, None] * Ny + j[:, None, :, None]) * Nx + i[:, None, None, :]
        return np.stack([(w * vslab[d].reshape(-1)[flat]).sum(axis=(1, 2, 3)) for d in range(3)])


class SlabRank:
    """One rank of the slab-decomposed transform.  ``exchange(blocks, shapes)``: blocks[s] goes to rank s, returns the
    list of blocks received from every rank, shapes[s] being the shape of the block rank s sends here (an all-to-all:
    NCCL grouped send/recv on the GPUs, gloo send/recv or plain lists in the tests)."""

    def __init__(self, model: PmeModel, R: int, r: int, exchange):
        self.m, self.R, self.r, self.exchange = model, R, r, exchange
        Nx, Ny, Nz = model.Nb
        if Nz % R:
            raise ValueError("Nb(3) must be a multiple of the rank count (SetEwaldPrms, ModConf.F90:394-395)")
        self.zs = slab_chunks(Nz, R)
        self.ys = slab_chunks(Ny, R)

    def forward(self, slab):
        """z-slab [comp][zloc][Ny][Nx] real -> y-slab [comp][Nz][yloc][Nxh] complex (Get_That)."""
        a = np.fft.rfft2(slab, axes=(-2, -1))                                                 # 2-D r2c per plane
        ylo, yhi = self.ys[self.r]
        shapes = [(a.shape[0], zhi - zlo, yhi - ylo, self.m.Nxh) for zlo, zhi in self.zs]
        got = self.exchange([np.ascontiguousarray(a[:, :, lo:hi, :]) for lo, hi in self.ys], shapes)
        b = np.concatenate(got, axis=1)                                                       # senders in rank order = z order
        return np.fft.ifft(b, axis=1) * self.m.Nb[2]                                          # z with e^{+i}

    def backward(self, yslab):
        """y-slab [3][Nz][yloc][Nxh] complex -> z-slab [3][zloc][Ny][Nx] real (Get_R)."""
        Nx, Ny, _ = self.m.Nb
        b = np.fft.fft(yslab, axis=1)                                                         # z with e^{-i}
        zlo, zhi = self.zs[self.r]
        shapes = [(b.shape[0], zhi - zlo, yhi - ylo, self.m.Nxh) for ylo, yhi in self.ys]
        got = self.exchange([np.ascontiguousarray(b[:, lo:hi]) for lo, hi in self.zs], shapes)
        a = np.concatenate(got, axis=2)                                                       # senders in rank order = y order
        return np.fft.irfft(np.fft.ifft(a, axis=2) * Ny, n=Nx, axis=3) * Nx                   # y with e^{+i}, c2r over x

    def transform(self, ff_slab, tt_slab):
        """PME_Transform on this rank's z-slab of the (already rank-summed) source meshes -> velocity mesh z-slab."""
        fC = self.forward(ff_slab) if ff_slab is not None else None
        tC = self.forward(tt_slab) if tt_slab is not None else None
        return self.backward(self.m.multiply(fC, tC, self.ys[self.r]))


def run_slabs_in_process(model: PmeModel, R: int, ff, tt):
    """All R ranks in one process (lock step through a list-based all-to-all): full source meshes in, full velocity
    mesh out -- the reference for what the distributed run must produce."""
    import threading
    box = [[None] * R for _ in range(R)]
    bar = threading.Barrier(R)

    def make_exchange(r):
        def exchange(blocks, shapes):
            for s in range(R):
                box[s][r] = blocks[s]
            bar.wait()
            got = list(box[r])
            assert all(g.shape == tuple(sh) for g, sh in zip(got, shapes))
            bar.wait()
            return got
        return exchange

    zs = slab_chunks(model.Nb[2], R)
    out = [None] * R

    def work(r):
        lo, hi = zs[r]
        rank = SlabRank(model, R, r, make_exchange(r))
        out[r] = rank.transform(None if ff is None else ff[:, lo:hi], None if tt is None else tt[:, lo:hi])

    th = [threading.Thread(target=work, args=(r,)) for r in range(R)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    return np.concatenate(out, axis=1)
