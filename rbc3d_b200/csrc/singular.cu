// singular.cu -- singular self-cell correction of the real-space sum: RBC_SingInt (ModRbcSingInt.F90:29-90),
// called from AddIntOnRbcs for every on-surface target (ModIntOnRbcs.F90:114-121).
//
// For a target (cell, ilat0, ilon0) the masked-out part of the pair sum is replaced by quadrature on a polar
// patch of nrad x nazm points whose reference-sphere coordinates thG/phiG are shared by all cells; positions,
// normals and densities at the patch points come from the cell's bicubic splines.
//
// v1: one warp per target, lanes stride over the patch points, splines gathered through the read-only path.
#include "device_math.cuh"
#include "rbc3d_internal.h"

namespace rbc3d {

constexpr int SING_WARPS = 8;

struct SingArgs {
  Params prm;
  int Np, npc, nlat, nlon, npatch;
  const double *th, *phi;
  const double *thG, *phiG, *pw;
  int nrad;
  const double *spx, *spa3, *spF, *spG;
  const double *Bcell;
  const int *active;
  const double *tab_sl, *tab_dl;
  double c1, c2;
  double *acc;  // SoA(3,Np)
};

template <bool SL, bool DL>
__global__ void __launch_bounds__(SING_WARPS * 32) k_singular(SingArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ti = blockIdx.x * SING_WARPS + warp;
  if (ti >= a.Np) return;
  if (!a.active[ti]) return;
  const int cell = ti / a.npc, pt = ti - cell * a.npc;
  const int ilon0 = pt / a.nlat, ilat0 = pt - ilon0 * a.nlat;
  const int m = 2 * a.nlat, n = a.nlon;
  const size_t sp3 = (size_t)12 * m * n;
  const double *spx = a.spx + sp3 * cell;
  const double *spa3 = a.spa3 + sp3 * cell;
  const double *spF = SL ? a.spF + sp3 * cell : nullptr;
  const double *spG = DL ? a.spG + sp3 * cell : nullptr;
  double xi[3];
  spline_interp<3>(spx, m, n, a.th[ilat0], a.phi[ilon0], xi);  // ModRbcSingInt.F90:58
  const double c2m = a.c2 * a.Bcell[cell];                      // c2Mod, ModIntOnRbcs.F90:116
  const size_t off = (size_t)pt * a.npatch;
  double dvx = 0, dvy = 0, dvz = 0;
  for (int q = lane; q < a.npatch; q += 32) {
    const double th_j = __ldg(a.thG + off + q), phi_j = __ldg(a.phiG + off + q);
    const double wq = __ldg(a.pw + (q % a.nrad));
    double xj[3];
    spline_interp<3>(spx, m, n, th_j, phi_j, xj);
    const double xx = xj[0] - xi[0], yy = xj[1] - xi[1], zz = xj[2] - xi[2];
    const double rr = sqrt(xx * xx + yy * yy + zz * zz);
    if (rr >= a.prm.rc) continue;
    if (SL) {
      double fj[3], EA, EB;
      spline_interp<3>(spF, m, n, th_j, phi_j, fj);
      fj[0] *= wq;
      fj[1] *= wq;
      fj[2] *= wq;
      ewald_sl(a.tab_sl, a.prm, rr, EA, EB);
      const double xf = EA * (xx * fj[0] + yy * fj[1] + zz * fj[2]);
      dvx += a.c1 * (xf * xx + EB * fj[0]);
      dvy += a.c1 * (xf * yy + EB * fj[1]);
      dvz += a.c1 * (xf * zz + EB * fj[2]);
    }
    if (DL) {
      double gj[3], nj[3];
      spline_interp<3>(spG, m, n, th_j, phi_j, gj);
      spline_interp<3>(spa3, m, n, th_j, phi_j, nj);
      const double EA = ewald_dl(a.tab_dl, a.prm, rr);
      const double qd = c2m * EA * wq * (xx * gj[0] + yy * gj[1] + zz * gj[2]) *
                        (xx * nj[0] + yy * nj[1] + zz * nj[2]);
      dvx += qd * xx;
      dvy += qd * yy;
      dvz += qd * zz;
    }
  }
  dvx = warp_sum(dvx);
  dvy = warp_sum(dvy);
  dvz = warp_sum(dvz);
  if (lane == 0) {
    a.acc[ti] += dvx;
    a.acc[(size_t)a.Np + ti] += dvy;
    a.acc[2 * (size_t)a.Np + ti] += dvz;
  }
}

int singular_prepare(rbc3d_ctx *) { return RBC3D_OK; }

int singular_apply(rbc3d_ctx *c, TargetList &t, double c1, double c2) {
  if (t.kind != RBC3D_TL_CELLS) return RBC3D_OK;  // only on-surface targets (ModIntOnRbcs.F90:115)
  Cells &C = c->cells;
  if (C.Np == 0) return RBC3D_OK;
  const bool sl = (c1 != 0), dl = (c2 != 0);
  if (!sl && !dl) return RBC3D_OK;
  SingArgs a;
  a.prm = c->prm;
  a.Np = C.Np;
  a.npc = C.npc;
  a.nlat = C.nlat;
  a.nlon = C.nlon;
  a.npatch = C.nrad * C.nazm;
  a.th = C.th.p;
  a.phi = C.phi.p;
  a.thG = C.thG.p;
  a.phiG = C.phiG.p;
  a.pw = C.pw.p;
  a.nrad = C.nrad;
  a.spx = C.spx.p;
  a.spa3 = C.spa3.p;
  a.spF = C.spF.p;
  a.spG = C.spG.p;
  a.Bcell = C.B.p;
  a.active = t.active.p;
  a.tab_sl = c->tab_sl.p;
  a.tab_dl = c->tab_dl.p;
  a.c1 = c1;
  a.c2 = c2;
  a.acc = t.acc.p;
  const int grid = (C.Np + SING_WARPS - 1) / SING_WARPS;
  if (sl && dl)
    k_singular<true, true><<<grid, SING_WARPS * 32, 0, c->stream>>>(a);
  else if (sl)
    k_singular<true, false><<<grid, SING_WARPS * 32, 0, c->stream>>>(a);
  else
    k_singular<false, true><<<grid, SING_WARPS * 32, 0, c->stream>>>(a);
  KERNEL_CHECK();
  c->launches++;
  return RBC3D_OK;
}

}  // namespace rbc3d
