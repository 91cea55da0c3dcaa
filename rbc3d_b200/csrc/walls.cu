// walls.cu -- wall surfaces of the Ewald boundary-integral operator (ModIntOnWalls.F90).
//
//  reference                                              here
//  SourceList_UpdateCoord(slist_wall, walls)              walls_set_geometry: element centroids + cell list (celllist.cu)
//    (ModSourceList.F90:127-146)
//  TargetList_Update(tlist_wall, walls)                   target list RBC3D_TL_WALLS: all wall vertices, Acoef = 2
//    (ModTargetList.F90:122-131)
//  AddIntOnWalls direct loop (ModIntOnWalls.F90:80-126)   k_wall_scan (geometry time: which (target, element) pairs are
//                                                         within rc, MinDistToTri, Duffy or 7-point rule) +
//                                                         k_wall_eval (per application) + k_wall_reduce
//  PrepareSingIntOnWall (ModIntOnWalls.F90:181-308)       the same scan restricted to the wall's own elements,
//                                                         k_wall_eval<LHS>, radix sort of the (row, col) block keys and
//                                                         a segmented sum: block-row sparse matrix in HBM
//  SingIntOnWall (ModIntOnWalls.F90:136-172)              k_wall_spmv (one warp per vertex row; HBM-bound)
//  PME wall sources (ModPME.F90:105-129)                  k_wall_pme_sources; spread by the PME kernels (pme.cu)
//
// A (target, element) pair is evaluated by 16 lanes: the 7 Gauss points of Tri_Int_Regular on lanes 0..6, or the
// 4 x 4 Gauss-Legendre points of each of the three Duffy sub-triangles (lane = point, loop over sub-triangles),
// followed by a fixed shuffle tree.  Every per-target / per-block sum runs in a fixed order: results are
// deterministic.  The range decision uses the reference's un-fused arithmetic (MinDistToTri) so that the pair sets
// are identical to the CPU code's.
#include <cub/cub.cuh>

#include "device_math.cuh"
#include "rbc3d_internal.h"

namespace rbc3d {

struct WallQuad {
  double tri_s[7], tri_t[7], tri_w[7];  // gqTri7, ModQuadRule.F90:70-94
  double gl_r[4], gl_w[4];              // GauLeg(0,1,4), ModIntOnWalls.F90:397-401
};
__constant__ WallQuad c_wq;

static int upload_quad(rbc3d_ctx *c) {
  static bool done[64] = {false};
  if (c->device < 64 && done[c->device]) return RBC3D_OK;
  WallQuad q;
  double r = (6. - sqrt(15.)) / 21., s = r, t = 1 - r - s;
  q.tri_s[0] = r, q.tri_s[1] = s, q.tri_s[2] = t;
  q.tri_t[0] = s, q.tri_t[1] = t, q.tri_t[2] = r;
  q.tri_w[0] = q.tri_w[1] = q.tri_w[2] = (155. - sqrt(15.)) / 2400.;
  r = (6. + sqrt(15.)) / 21., s = r, t = 1 - r - s;
  q.tri_s[3] = r, q.tri_s[4] = s, q.tri_s[5] = t;
  q.tri_t[3] = s, q.tri_t[4] = t, q.tri_t[5] = r;
  q.tri_w[3] = q.tri_w[4] = q.tri_w[5] = (155. + sqrt(15.)) / 2400.;
  q.tri_s[6] = 1. / 3., q.tri_t[6] = 1. / 3.;
  q.tri_w[6] = 9. / 80.;
  h_gauleg(0., 1., 4, q.gl_r, q.gl_w);
  CUDA_TRY(cudaMemcpyToSymbolAsync(c_wq, &q, sizeof(q), 0, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (c->device < 64) done[c->device] = true;
  return RBC3D_OK;
}

// ---------------------------------------------------------------------------------------------------------
// un-fused helpers for the range decision
__device__ __forceinline__ double mul_(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dot3_(const double *a, const double *b) {
  return add_(add_(mul_(a[0], b[0]), mul_(a[1], b[1])), mul_(a[2], b[2]));
}

// ModIntOnWalls.F90:480-577 MinDistToTri; x[l][d]
__device__ __forceinline__ double min_dist_to_tri(const double xTar[3], const double x[3][3], double &s0, double &t0) {
  double x12[3], x13[3], x1Tar[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    x12[k] = sub_(x[1][k], x[0][k]);
    x13[k] = sub_(x[2][k], x[0][k]);
    x1Tar[k] = sub_(x[0][k], xTar[k]);
  }
  const double a = dot3_(x12, x12), b = dot3_(x12, x13), c = dot3_(x13, x13);
  const double det = sub_(mul_(a, c), mul_(b, b));
  const double invDet = __ddiv_rn(1., det);
  const double d = dot3_(x12, x1Tar), e = dot3_(x13, x1Tar), f = dot3_(x1Tar, x1Tar);
  double s = sub_(mul_(b, e), mul_(c, d));
  double t = sub_(mul_(b, d), mul_(a, e));
  int region;
  if (add_(s, t) <= det) {
    if (s < 0)
      region = (t < 0) ? 4 : 3;
    else if (t < 0)
      region = 5;
    else
      region = 0;
  } else {
    if (s < 0)
      region = 2;
    else if (t < 0)
      region = 6;
    else
      region = 1;
  }
  if (region == 2)
    region = (-add_(c, e) < 0) ? 3 : 1;
  else if (region == 4)
    region = (d < 0) ? 5 : 3;
  else if (region == 6)
    region = (sub_(sub_(add_(b, e), a), d) < 0) ? 1 : 5;
  switch (region) {
    case 0:
      s = mul_(invDet, s);
      t = mul_(invDet, t);
      break;
    case 1:
      s = __ddiv_rn(sub_(sub_(add_(c, e), b), d), add_(sub_(a, mul_(2., b)), c));
      s = fmin(1., fmax(0., s));
      t = sub_(1., s);
      break;
    case 3:
      s = 0.;
      t = __ddiv_rn(-e, c);
      t = fmin(1., fmax(0., t));
      break;
    default:
      t = 0.;
      s = __ddiv_rn(-d, a);
      s = fmin(1., fmax(0., s));
      break;
  }
  // a*s*s + 2*b*s*t + c*t*t + 2*d*s + 2*e*t + f, left to right
  double q = mul_(mul_(a, s), s);
  q = add_(q, mul_(mul_(mul_(2., b), s), t));
  q = add_(q, mul_(mul_(c, t), t));
  q = add_(q, mul_(mul_(2., d), s));
  q = add_(q, mul_(mul_(2., e), t));
  q = add_(q, f);
  s0 = s;
  t0 = t;
  return __dsqrt_rn(q);
}

struct WallGeom {
  int NV, NE;
  const double *x;     // SoA(3,NV)
  const int *e2v;      // SoA(3,NE) global 0-based
  const int *ewall;    // [NE]
  const double *epsDist;
};

// element e translated close to xi (ModIntOnWalls.F90:96-107): xx = nint((xi - xele(1,:))*iLb)*Lb
__device__ __forceinline__ void load_element(const WallGeom &g, const Params &prm, int e, const double xi[3],
                                             double x[3][3], int iv[3]) {
#pragma unroll
  for (int l = 0; l < 3; l++) {
    iv[l] = g.e2v[(size_t)l * g.NE + e];
#pragma unroll
    for (int d = 0; d < 3; d++) x[l][d] = g.x[(size_t)d * g.NV + iv[l]];
  }
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const double sh = mul_(round(mul_(sub_(xi[d], x[0][d]), prm.iLb[d])), prm.Lb[d]);
#pragma unroll
    for (int l = 0; l < 3; l++) x[l][d] = add_(x[l][d], sh);
  }
}

// ---------------------------------------------------------------------------------------------------------
// geometry time: (target, element) pairs within rc.  MODE 0: skip elements of the target's own surface
// (AddIntOnWalls, :92); MODE 1: only elements of the target's own wall (PrepareSingIntOnWall, :263).
struct ScanArgs {
  Params prm;
  WallGeom g;
  int n;                    // targets
  const double *tx;         // SoA(3,n)
  const int *tactive;
  const int *tsurf;         // wall index of a wall-vertex target, < 0 otherwise (cells: never equal to a wall)
  const int *tcid;          // real-space cell of the target (ncells = inactive)
  const int *wstart, *worder;  // cell list of the element centroids
  const int *off;           // FILL: exclusive offsets
  int *cnt;                 // COUNT
  int *ele, *kind;          // FILL: element, 1 = Duffy
  double *s0, *t0;
  int *ptarget;             // FILL: target of the pair
};

template <int MODE, bool FILL>
__global__ void __launch_bounds__(128) k_wall_scan(ScanArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  int n = 0;
  const int pos0 = FILL ? a.off[i] : 0;
  if (a.tactive[i]) {
    const double xi[3] = {a.tx[i], a.tx[(size_t)a.n + i], a.tx[2 * (size_t)a.n + i]};
    const int *Nc = a.prm.Nc;
    const int cid = a.tcid[i];
    const int i1 = cid % Nc[0], i2 = (cid / Nc[0]) % Nc[1], i3 = cid / (Nc[0] * Nc[1]);
    const int surf = a.tsurf[i];
    for (int d1 = -1; d1 <= 1; d1++)
      for (int d2 = -1; d2 <= 1; d2++)
        for (int d3 = -1; d3 <= 1; d3++) {
          const int j1 = imodulo(i1 + d1, Nc[0]), j2 = imodulo(i2 + d2, Nc[1]), j3 = imodulo(i3 + d3, Nc[2]);
          const int cj = j1 + Nc[0] * (j2 + Nc[1] * j3);
          for (int k = a.wstart[cj]; k < a.wstart[cj + 1]; k++) {
            const int e = a.worder[k];
            const int w = a.g.ewall[e];
            if (MODE == 0 ? (surf == w) : (surf != w)) continue;
            double x[3][3], s0, t0;
            int iv[3];
            load_element(a.g, a.prm, e, xi, x, iv);
            const double rr = min_dist_to_tri(xi, x, s0, t0);
            if (rr > a.prm.rc) continue;
            if (FILL) {
              const int p = pos0 + n;
              a.ele[p] = e;
              // AddIntOnWalls: Duffy if rr < epsDist (:112); PrepareSingIntOnWall: regular if rr > epsDist (:283)
              a.kind[p] = (MODE == 0) ? (rr < a.g.epsDist[e]) : !(rr > a.g.epsDist[e]);
              a.s0[p] = s0;
              a.t0[p] = t0;
              a.ptarget[p] = i;
            }
            n++;
          }
        }
  }
  if (!FILL) a.cnt[i] = n;
}

// ---------------------------------------------------------------------------------------------------------
// one quadrature point: (EA xx xx^T + EB I) dsGq, ModIntOnWalls.F90:344-360 / 424-458
__device__ __forceinline__ void sl_point(const Params &prm, const double *__restrict__ tab_sl, const double xtar[3],
                                         const double xGq[3], double &EA, double &EB, double xx[3]) {
#pragma unroll
  for (int k = 0; k < 3; k++) xx[k] = xtar[k] - xGq[k];
  const double rr = sqrt(xx[0] * xx[0] + xx[1] * xx[1] + xx[2] * xx[2]);
  ewald_sl(tab_sl, prm, rr, EA, EB);
}

// sum over the 16 lanes of a pair (the other half of the warp may work on a different branch or have left)
__device__ __forceinline__ double group_sum16(double v) {
  const unsigned mask = (threadIdx.x & 16) ? 0xffff0000u : 0x0000ffffu;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, 16);
  return v;
}

struct EvalArgs {
  Params prm;
  WallGeom g;
  int npair;
  const int *ptarget, *ele, *kind;
  const double *s0, *t0;
  const double *tx;    // target coordinates SoA(3,n)
  int n;
  const double *f;     // SoA(3,NV) tractions (RHS mode)
  const double *tab_sl;
  double *dv;          // RHS: SoA(3,npair)
  double *lhs;         // LHS: [npair][27]  lhs(l,ii,jj)
};

// 16 lanes per pair
template <bool LHS>
__global__ void __launch_bounds__(128) k_wall_eval(EvalArgs a) {
  const int gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const int lane = threadIdx.x & 15;
  if (gid >= a.npair) return;  // whole 16-lane group leaves together
  const int i = a.ptarget[gid], e = a.ele[gid];
  const double xi[3] = {a.tx[i], a.tx[(size_t)a.n + i], a.tx[2 * (size_t)a.n + i]};
  double x[3][3], f[3][3];
  int iv[3];
  load_element(a.g, a.prm, e, xi, x, iv);
#pragma unroll
  for (int l = 0; l < 3; l++)
#pragma unroll
    for (int d = 0; d < 3; d++) f[l][d] = LHS ? 0.0 : a.f[(size_t)d * a.g.NV + iv[l]];
  double rhs[3] = {0., 0., 0.};
  double lhs[27];
  if (LHS) {
#pragma unroll
    for (int q = 0; q < 27; q++) lhs[q] = 0.;
  }
  auto accum = [&](const double xGq[3], const double fG[3], double dsGq, double w0, double w1, double w2) {
    double EA, EB, xx[3];
    sl_point(a.prm, a.tab_sl, xi, xGq, EA, EB, xx);
    if (!LHS) {
      const double fq[3] = {dsGq * fG[0], dsGq * fG[1], dsGq * fG[2]};
      const double dot = xx[0] * fq[0] + xx[1] * fq[1] + xx[2] * fq[2];
#pragma unroll
      for (int k = 0; k < 3; k++) rhs[k] += EA * xx[k] * dot + EB * fq[k];
    } else {
#pragma unroll
      for (int ii = 0; ii < 3; ii++)
#pragma unroll
        for (int jj = 0; jj < 3; jj++) {
          double v = EA * xx[ii] * xx[jj];
          if (ii == jj) v += EB;
          v *= dsGq;
          lhs[ii * 3 + jj] += w0 * v;
          lhs[9 + ii * 3 + jj] += w1 * v;
          lhs[18 + ii * 3 + jj] += w2 * v;
        }
    }
  };
  if (!a.kind[gid]) {
    // Tri_Int_Regular, :319-363
    if (lane < 7) {
      double c[3];
      {
        double u[3], v[3];
#pragma unroll
        for (int d = 0; d < 3; d++) u[d] = x[1][d] - x[0][d], v[d] = x[2][d] - x[0][d];
        c[0] = u[1] * v[2] - u[2] * v[1];
        c[1] = u[2] * v[0] - u[0] * v[2];
        c[2] = u[0] * v[1] - u[1] * v[0];
      }
      const double detJ = 2 * (0.5 * sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]));
      const double s = c_wq.tri_s[lane], t = c_wq.tri_t[lane];
      double xGq[3], fGq[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        xGq[k] = (1. - s - t) * x[0][k] + s * x[1][k] + t * x[2][k];
        fGq[k] = (1. - s - t) * f[0][k] + s * f[1][k] + t * f[2][k];
      }
      accum(xGq, fGq, c_wq.tri_w[lane] * detJ, 1 - s - t, s, t);
    }
  } else {
    // Tri_Int_Duffy, :373-465: lane = (i, j) of the 4 x 4 rule, loop over the three sub-triangles
    const double s0 = a.s0[gid], t0 = a.t0[gid];
    double x0[3], f0[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      x0[k] = (1 - s0 - t0) * x[0][k] + s0 * x[1][k] + t0 * x[2][k];
      f0[k] = (1 - s0 - t0) * f[0][k] + s0 * f[1][k] + t0 * f[2][k];
    }
    const int qi = lane >> 2, qj = lane & 3;
    const double s = c_wq.gl_r[qi], t = s * c_wq.gl_r[qj];
    const double wq = c_wq.gl_w[qi] * c_wq.gl_w[qj];
#pragma unroll
    for (int n = 0; n < 3; n++) {
      const int n2 = (n + 1) % 3;
      double u[3], v[3], nr[3];
#pragma unroll
      for (int k = 0; k < 3; k++) u[k] = x[n][k] - x0[k], v[k] = x[n2][k] - x0[k];
      nr[0] = u[1] * v[2] - u[2] * v[1];
      nr[1] = u[2] * v[0] - u[0] * v[2];
      nr[2] = u[0] * v[1] - u[1] * v[0];
      const double detJ = sqrt(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]);
      const double s1 = (n == 1) ? 1. : 0., t1 = (n == 2) ? 1. : 0.;
      const double s2 = (n == 0) ? 1. : 0., t2 = (n == 1) ? 1. : 0.;
      double xGq[3], fGq[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        xGq[k] = (1. - s) * x0[k] + (s - t) * x[n][k] + t * x[n2][k];
        fGq[k] = (1. - s) * f0[k] + (s - t) * f[n][k] + t * f[n2][k];
      }
      const double sG = (1. - s) * s0 + (s - t) * s1 + t * s2;
      const double tG = (1. - s) * t0 + (s - t) * t1 + t * t2;
      accum(xGq, fGq, wq * detJ * s, 1. - sG - tG, sG, tG);
    }
  }
  if (!LHS) {
#pragma unroll
    for (int k = 0; k < 3; k++) rhs[k] = group_sum16(rhs[k]);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 3; k++) a.dv[(size_t)k * a.npair + gid] = rhs[k];
    }
  } else {
#pragma unroll
    for (int q = 0; q < 27; q++) lhs[q] = group_sum16(lhs[q]);
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < 27; q++) a.lhs[(size_t)gid * 27 + q] = lhs[q];
    }
  }
}

// acc(i,:) += c1 * sum of the target's pairs, in list order (v(i,:) += c1*dv/Acoef, :117; /Acoef in combine())
__global__ void k_wall_reduce(int n, const int *__restrict__ off, int npair, const double *__restrict__ dv, double c1,
                              double *__restrict__ acc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = off[i], e = off[i + 1];
  if (b == e) return;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double s = 0.;
    for (int p = b; p < e; p++) s += dv[(size_t)k * npair + p];
    acc[(size_t)k * n + i] += c1 * s;
  }
}

// ---------------------------------------------------------------------------------------------------------
// matrix assembly: keys (row vertex, column vertex) of the 3 blocks of every pair
__global__ void k_wall_keys(int npair, const int *__restrict__ ptarget, const int *__restrict__ ele, int NE,
                            const int *__restrict__ e2v, unsigned long long *__restrict__ keys, int *__restrict__ vals) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= 3 * npair) return;
  const int p = q / 3, l = q - 3 * p;
  const unsigned long long row = (unsigned)ptarget[p], col = (unsigned)e2v[(size_t)l * NE + ele[p]];
  keys[q] = (row << 32) | col;
  vals[q] = q;
}
__global__ void k_wall_heads(int m, const unsigned long long *__restrict__ keys, int *__restrict__ head) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= m) return;
  head[q] = (q == 0 || keys[q] != keys[q - 1]) ? 1 : 0;
}
// one thread per sorted entry that starts a block: sum the run (ascending pair order = MatSetValues ADD order up to
// the traversal order of the linked list), write the block, its column, and count it for its row
__global__ void k_wall_blocks(int m, const unsigned long long *__restrict__ keys, const int *__restrict__ vals,
                              const int *__restrict__ head, const int *__restrict__ bidx /* exclusive scan of head */,
                              const double *__restrict__ lhs, int *__restrict__ col, double *__restrict__ val,
                              int *__restrict__ rowcnt) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= m || !head[q]) return;
  const unsigned long long key = keys[q];
  double s[9];
#pragma unroll
  for (int k = 0; k < 9; k++) s[k] = 0.;
  for (int r = q; r < m && keys[r] == key; r++) {
    const int ent = vals[r], p = ent / 3, l = ent - 3 * p;
    const double *src = lhs + (size_t)p * 27 + 9 * l;
#pragma unroll
    for (int k = 0; k < 9; k++) s[k] += src[k];
  }
  const int b = bidx[q];
  col[b] = (int)(key & 0xffffffffu);
#pragma unroll
  for (int k = 0; k < 9; k++) val[(size_t)b * 9 + k] = s[k];
  atomicAdd(&rowcnt[(int)(key >> 32)], 1);
}

// SingIntOnWall + the accumulation of AddIntOnWalls :61-66: out(i,:) (+)= c1 * sum_blocks val * f(col).
// One warp per vertex row, lanes stride over the row's blocks, fixed shuffle tree.
__global__ void __launch_bounds__(256) k_wall_spmv(int v_lo, int v_hi, int NV, const int *__restrict__ rowptr,
                                                   const int *__restrict__ col, const double *__restrict__ val,
                                                   const double *__restrict__ f, double c1, double *__restrict__ out,
                                                   int out_n, int out_off, int accumulate) {
  const int row = v_lo + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= v_hi) return;
  double s0 = 0., s1 = 0., s2 = 0.;
  for (int b = rowptr[row] + lane; b < rowptr[row + 1]; b += 32) {
    const int cv = col[b];
    const double f0 = f[cv], f1 = f[(size_t)NV + cv], f2 = f[2 * (size_t)NV + cv];
    const double *m = val + (size_t)b * 9;
    s0 += m[0] * f0 + m[1] * f1 + m[2] * f2;
    s1 += m[3] * f0 + m[4] * f1 + m[5] * f2;
    s2 += m[6] * f0 + m[7] * f1 + m[8] * f2;
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) {
    const size_t o = (size_t)(row - out_off);
    if (accumulate) {
      out[o] += c1 * s0;
      out[(size_t)out_n + o] += c1 * s1;
      out[2 * (size_t)out_n + o] += c1 * s2;
    } else {
      out[o] = c1 * s0;
      out[(size_t)out_n + o] = c1 * s1;
      out[2 * (size_t)out_n + o] = c1 * s2;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
__global__ void k_wall_centroids(int NE, int NV, const double *__restrict__ x, const int *__restrict__ e2v,
                                 double *__restrict__ xc) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= NE) return;
  const int a = e2v[e], b = e2v[(size_t)NE + e], c = e2v[2 * (size_t)NE + e];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const double *xd = x + (size_t)d * NV;
    // THRD*sum(xele, dim=1), ModSourceList.F90:139
    xc[(size_t)d * NE + e] = mul_(1.0 / 3, add_(add_(xd[a], xd[b]), xd[c]));
  }
}
// ftmp = THRD*sum(fele,dim=1)*area, ModPME.F90:119
__global__ void k_wall_pme_sources(int NE, int NV, const double *__restrict__ f, const int *__restrict__ e2v,
                                   const double *__restrict__ area, double *__restrict__ ft) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= NE) return;
  const int a = e2v[e], b = e2v[(size_t)NE + e], c = e2v[2 * (size_t)NE + e];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const double *fd = f + (size_t)d * NV;
    ft[(size_t)d * NE + e] = (1.0 / 3) * ((fd[a] + fd[b]) + fd[c]) * area[e];
  }
}
__global__ void k_wall_target_meta(int NV, const int *__restrict__ vwall, int *__restrict__ surf,
                                   double *__restrict__ A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NV) return;
  (void)vwall;
  surf[i] = -1;  // never a cell surface (the pair / near-singular kernels compare surf with cell indices)
  A[i] = 2.0;    // TargetList_Update, ModTargetList.F90:127
}

// ---------------------------------------------------------------------------------------------------------
static WallGeom wall_geom(const Walls &W) {
  WallGeom g;
  g.NV = W.NV;
  g.NE = W.NE;
  g.x = W.x.p;
  g.e2v = W.e2v.p;
  g.ewall = W.ewall.p;
  g.epsDist = W.epsDist.p;
  return g;
}

int walls_set_geometry(rbc3d_ctx *c, int nwall, const int *nvert, const int *nele, const double *x, const int *e2v,
                       const double *area, const double *epsDist) {
  Walls &W = c->walls;
  RBC_TRY(upload_quad(c));
  W.nwall = nwall;
  W.h_nvert.assign(nvert, nvert + nwall);
  W.h_nele.assign(nele, nele + nwall);
  W.h_voff.assign(nwall + 1, 0);
  W.h_eoff.assign(nwall + 1, 0);
  for (int w = 0; w < nwall; w++) {
    W.h_voff[w + 1] = W.h_voff[w] + nvert[w];
    W.h_eoff[w + 1] = W.h_eoff[w] + nele[w];
  }
  const int NV = W.NV = W.h_voff[nwall], NE = W.NE = W.h_eoff[nwall];
  W.mat_ok = false;
  W.f_set = false;
  W.geom_version++;
  const size_t nv1 = NV > 0 ? NV : 1, ne1 = NE > 0 ? NE : 1;
  std::vector<int> e2g(3 * ne1), ewall(ne1), vwall(nv1);
  for (int w = 0; w < nwall; w++) {
    for (int e = W.h_eoff[w]; e < W.h_eoff[w + 1]; e++) {
      ewall[e] = w;
      for (int l = 0; l < 3; l++) {
        const int loc = e2v[(size_t)l * NE + e];
        if (loc < 1 || loc > nvert[w]) {
          set_error("rbc3d_walls_set: e2v(%d,%d) = %d outside 1..%d", e - W.h_eoff[w] + 1, l + 1, loc, nvert[w]);
          return RBC3D_EINVAL;
        }
        e2g[(size_t)l * NE + e] = W.h_voff[w] + loc - 1;
      }
    }
    for (int v = W.h_voff[w]; v < W.h_voff[w + 1]; v++) vwall[v] = w;
  }
  RBC_TRY(W.x.resize(3 * nv1));
  RBC_TRY(W.f.resize(3 * nv1));
  RBC_TRY(W.e2v.resize(3 * ne1));
  RBC_TRY(W.ewall.resize(ne1));
  RBC_TRY(W.vwall.resize(nv1));
  RBC_TRY(W.area.resize(ne1));
  RBC_TRY(W.epsDist.resize(ne1));
  RBC_TRY(W.xc.resize(3 * ne1));
  RBC_TRY(W.ft.resize(3 * ne1));
  if (NV > 0) {
    CUDA_TRY(cudaMemcpyAsync(W.x.p, x, sizeof(double) * 3 * NV, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(W.vwall.p, vwall.data(), sizeof(int) * NV, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemsetAsync(W.f.p, 0, sizeof(double) * 3 * NV, c->stream));
  }
  if (NE > 0) {
    CUDA_TRY(cudaMemcpyAsync(W.e2v.p, e2g.data(), sizeof(int) * 3 * NE, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(W.ewall.p, ewall.data(), sizeof(int) * NE, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(W.area.p, area, sizeof(double) * NE, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(W.epsDist.p, epsDist, sizeof(double) * NE, cudaMemcpyHostToDevice, c->stream));
    k_wall_centroids<<<(NE + 255) / 256, 256, 0, c->stream>>>(NE, NV, W.x.p, W.e2v.p, W.xc.p);
    KERNEL_CHECK();
    c->launches++;
    CUDA_TRY(cudaMemsetAsync(W.ft.p, 0, sizeof(double) * 3 * NE, c->stream));
  }
  CUDA_TRY(cudaStreamSynchronize(c->stream));  // host staging vectors go out of scope
  // slist_wall: HashTable_Build over the centroids; PME blocks for spreading (every rank spreads a slice)
  RBC_TRY(celllist_build_realspace(c, W.cl, NE, W.xc.p, nullptr));
  const int *own = nullptr;
  if (c->prm.nranks > 1 && NE > 0) {
    // element centroids whose B-spline support touches this rank's z-slab of mesh planes (pme.cu), or an index block
    RBC_TRY(pme_source_ownership(c, NE, W.xc.p, W.src_own, &own));
    if (!own) {
      std::vector<int> o(NE, 0);
      const int lo = (int)((long long)NE * c->prm.rank / c->prm.nranks);
      const int hi = (int)((long long)NE * (c->prm.rank + 1) / c->prm.nranks);
      for (int e = lo; e < hi; e++) o[e] = 1;
      CUDA_TRY(cudaMemcpy(W.src_own.p, o.data(), sizeof(int) * NE, cudaMemcpyHostToDevice));
      own = W.src_own.p;
    }
  }
  pme_spread_mode(c, W.pl, NE);
  RBC_TRY(celllist_build_pme(c, W.pl, NE, W.xc.p, own, W.pl.sblk, W.pl.swalk));
  if (W.pl.swalk) RBC_TRY(celllist_pme_weights(c, W.pl, W.xc.p));
  W.geom_set = true;
  return RBC3D_OK;
}

int walls_set_traction(rbc3d_ctx *c, const double *f, bool from_device) {
  Walls &W = c->walls;
  if (!W.geom_set) return RBC3D_ESTATE;
  if (W.NV > 0 && f != W.f.p)
    CUDA_TRY(cudaMemcpyAsync(W.f.p, f, sizeof(double) * 3 * W.NV, from_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                             c->stream));
  if (W.NE > 0) {
    k_wall_pme_sources<<<(W.NE + 255) / 256, 256, 0, c->stream>>>(W.NE, W.NV, W.f.p, W.e2v.p, W.area.p, W.ft.p);
    KERNEL_CHECK();
    c->launches++;
  }
  W.f_set = true;
  return RBC3D_OK;
}

int walls_target_meta(rbc3d_ctx *c, TargetList &t) {
  Walls &W = c->walls;
  if (W.NV > 0) {
    k_wall_target_meta<<<(W.NV + 255) / 256, 256, 0, c->stream>>>(W.NV, W.vwall.p, t.surf.p, t.Acoef.p);
    KERNEL_CHECK();
  }
  return RBC3D_OK;
}

// (target, element) pair list of a target list; mode 0 = AddIntOnWalls (other surfaces), 1 = own wall
static int scan_pairs(rbc3d_ctx *c, TargetList &t, WallPairs &wp, int mode, const int *tsurf) {
  Walls &W = c->walls;
  wp.npair = 0;
  wp.valid = false;
  const int n = t.n;
  RBC_TRY(wp.off.resize((size_t)n + 2));
  CUDA_TRY(cudaMemsetAsync(wp.off.p, 0, sizeof(int) * ((size_t)n + 2), c->stream));
  if (n == 0 || W.NE == 0) {
    wp.valid = true;
    return RBC3D_OK;
  }
  ScanArgs a;
  a.prm = c->prm;
  a.g = wall_geom(W);
  a.n = n;
  a.tx = t.x.p;
  a.tactive = t.active.p;
  a.tsurf = tsurf;
  a.tcid = t.cl.cid.p;
  a.wstart = W.cl.start.p;
  a.worder = W.cl.order.p;
  a.off = nullptr;
  a.cnt = wp.off.p;
  a.ele = a.kind = a.ptarget = nullptr;
  a.s0 = a.t0 = nullptr;
  const int grid = (n + 127) / 128;
  if (mode == 0)
    k_wall_scan<0, false><<<grid, 128, 0, c->stream>>>(a);
  else
    k_wall_scan<1, false><<<grid, 128, 0, c->stream>>>(a);
  KERNEL_CHECK();
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, wp.off.p, wp.off.p, n + 1, c->stream);
  RBC_TRY(wp.tmp.resize(bytes + 256));
  size_t avail = wp.tmp.n;
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(wp.tmp.p, avail, wp.off.p, wp.off.p, n + 1, c->stream));
  int np = 0;
  CUDA_TRY(cudaMemcpyAsync(&np, wp.off.p + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  wp.npair = np;
  const size_t np1 = np > 0 ? np : 1;
  RBC_TRY(wp.ele.resize(np1));
  RBC_TRY(wp.kind.resize(np1));
  RBC_TRY(wp.ptarget.resize(np1));
  RBC_TRY(wp.s0.resize(np1));
  RBC_TRY(wp.t0.resize(np1));
  RBC_TRY(wp.dv.resize(3 * np1));
  if (np > 0) {
    a.off = wp.off.p;
    a.cnt = nullptr;
    a.ele = wp.ele.p;
    a.kind = wp.kind.p;
    a.ptarget = wp.ptarget.p;
    a.s0 = wp.s0.p;
    a.t0 = wp.t0.p;
    if (mode == 0)
      k_wall_scan<0, true><<<grid, 128, 0, c->stream>>>(a);
    else
      k_wall_scan<1, true><<<grid, 128, 0, c->stream>>>(a);
    KERNEL_CHECK();
  }
  c->launches += 3;
  wp.valid = true;
  wp.geom_version = W.geom_version;
  wp.tl_version = t.version;
  return RBC3D_OK;
}

// wall index per target for the same-surface tests: wall targets carry it in t.surf, every other list -1
static int wall_surf(rbc3d_ctx *c, TargetList &t, const int **out) {
  Walls &W = c->walls;
  if (t.kind == RBC3D_TL_WALLS) {
    *out = W.vwall.p;
    return RBC3D_OK;
  }
  const size_t n1 = t.n > 0 ? t.n : 1;
  if (W.minus1.n < n1) {
    RBC_TRY(W.minus1.resize(n1));
    CUDA_TRY(cudaMemsetAsync(W.minus1.p, 0xff, sizeof(int) * n1, c->stream));
  }
  *out = W.minus1.p;
  return RBC3D_OK;
}

static EvalArgs eval_args(rbc3d_ctx *c, TargetList &t, WallPairs &wp) {
  Walls &W = c->walls;
  EvalArgs a;
  a.prm = c->prm;
  a.g = wall_geom(W);
  a.npair = wp.npair;
  a.ptarget = wp.ptarget.p;
  a.ele = wp.ele.p;
  a.kind = wp.kind.p;
  a.s0 = wp.s0.p;
  a.t0 = wp.t0.p;
  a.tx = t.x.p;
  a.n = t.n;
  a.f = W.f.p;
  a.tab_sl = c->tab_sl.p;
  a.dv = wp.dv.p;
  a.lhs = nullptr;
  return a;
}

// PrepareSingIntOnWall for every wall (rows of the ACTIVE wall vertices, :199-203)
int walls_prepare_sing(rbc3d_ctx *c) {
  Walls &W = c->walls;
  TargetList &t = c->tl[RBC3D_TL_WALLS];
  if (!W.geom_set || !t.valid) {
    set_error("PrepareSingIntOnWall: walls not set");
    return RBC3D_ESTATE;
  }
  W.mat_ok = false;
  const int NV = W.NV;
  WallPairs wp;
  RBC_TRY(scan_pairs(c, t, wp, 1, W.vwall.p));
  const int np = wp.npair, m = 3 * np;
  RBC_TRY(W.rowptr.resize((size_t)NV + 2));
  CUDA_TRY(cudaMemsetAsync(W.rowptr.p, 0, sizeof(int) * ((size_t)NV + 2), c->stream));
  W.nblk = 0;
  if (np > 0) {
    dbuf<double> lhs;
    dbuf<unsigned long long> keys, keys2;
    dbuf<int> vals, vals2, head, bidx;
    dbuf<char> tmp;
    RBC_TRY(lhs.resize((size_t)np * 27));
    RBC_TRY(keys.resize(m));
    RBC_TRY(keys2.resize(m));
    RBC_TRY(vals.resize(m));
    RBC_TRY(vals2.resize(m));
    RBC_TRY(head.resize((size_t)m + 1));
    RBC_TRY(bidx.resize((size_t)m + 1));
    EvalArgs a = eval_args(c, t, wp);
    a.lhs = lhs.p;
    k_wall_eval<true><<<(int)(((size_t)np * 16 + 127) / 128), 128, 0, c->stream>>>(a);
    KERNEL_CHECK();
    k_wall_keys<<<(m + 255) / 256, 256, 0, c->stream>>>(np, wp.ptarget.p, wp.ele.p, W.NE, W.e2v.p, keys.p, vals.p);
    KERNEL_CHECK();
    size_t b1 = 0, b2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, keys.p, keys2.p, vals.p, vals2.p, m, 0, 64, c->stream);
    cub::DeviceScan::ExclusiveSum(nullptr, b2, head.p, bidx.p, m + 1, c->stream);
    RBC_TRY(tmp.resize((b1 > b2 ? b1 : b2) + 256));
    size_t avail = tmp.n;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp.p, avail, keys.p, keys2.p, vals.p, vals2.p, m, 0, 64, c->stream));
    CUDA_TRY(cudaMemsetAsync(head.p, 0, sizeof(int) * ((size_t)m + 1), c->stream));
    k_wall_heads<<<(m + 255) / 256, 256, 0, c->stream>>>(m, keys2.p, head.p);
    KERNEL_CHECK();
    avail = tmp.n;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, avail, head.p, bidx.p, m + 1, c->stream));
    int nblk = 0;
    CUDA_TRY(cudaMemcpyAsync(&nblk, bidx.p + m, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    W.nblk = nblk;
    RBC_TRY(W.col.resize(nblk));
    RBC_TRY(W.val.resize((size_t)nblk * 9));
    k_wall_blocks<<<(m + 255) / 256, 256, 0, c->stream>>>(m, keys2.p, vals2.p, head.p, bidx.p, lhs.p, W.col.p,
                                                          W.val.p, W.rowptr.p);
    KERNEL_CHECK();
    avail = tmp.n;
    // rowptr = exclusive scan of the per-row block counts (blocks are already sorted by row, then column)
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, avail, W.rowptr.p, W.rowptr.p, NV + 1, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->launches += 7;
    lhs.release(), keys.release(), keys2.release(), vals.release(), vals2.release(), head.release(), bidx.release();
    tmp.release();
  }
  wp.release();
  W.mat_ok = true;
  W.mat_version = W.geom_version;
  return RBC3D_OK;
}

// v(0:nvert) = c1 * lhs * f for wall iwall, into a device SoA(3,nvert) buffer
int walls_sing_int(rbc3d_ctx *c, double c1, int iwall, double *v_dev) {
  Walls &W = c->walls;
  if (!W.mat_ok || W.mat_version != W.geom_version) {
    set_error("SingIntOnWall: PrepareSingIntOnWall has not been called for this wall geometry");
    return RBC3D_ESTATE;
  }
  const int lo = W.h_voff[iwall], hi = W.h_voff[iwall + 1];
  if (hi > lo) {
    k_wall_spmv<<<((hi - lo) * 32 + 255) / 256, 256, 0, c->stream>>>(lo, hi, W.NV, W.rowptr.p, W.col.p, W.val.p,
                                                                     W.f.p, c1, v_dev, hi - lo, lo, 0);
    KERNEL_CHECK();
    c->launches++;
  }
  return RBC3D_OK;
}

// AddIntOnWalls(c1, tlist, v): un-normalised sums into t.acc
int walls_add_int(rbc3d_ctx *c, TargetList &t, double c1) {
  Walls &W = c->walls;
  if (W.nwall == 0 || !W.geom_set) return RBC3D_OK;
  if (!W.f_set) {
    set_error("AddIntOnWalls: wall tractions not set");
    return RBC3D_ESTATE;
  }
  if (t.kind == RBC3D_TL_WALLS) {  // self-interactions, :54-77
    if (!W.mat_ok || W.mat_version != W.geom_version) {
      set_error("AddIntOnWalls on wall targets: PrepareSingIntOnWall has not been called");
      return RBC3D_ESTATE;
    }
    if (W.NV > 0) {
      k_wall_spmv<<<(W.NV * 32 + 255) / 256, 256, 0, c->stream>>>(0, W.NV, W.NV, W.rowptr.p, W.col.p, W.val.p, W.f.p,
                                                                  c1, t.acc.p, t.n, 0, 1);
      KERNEL_CHECK();
      c->launches++;
    }
    if (W.nwall == 1) return RBC3D_OK;
  }
  WallPairs &wp = t.wp;
  if (!wp.valid || wp.geom_version != W.geom_version || wp.tl_version != t.version) {
    const int *surf = nullptr;
    RBC_TRY(wall_surf(c, t, &surf));
    RBC_TRY(scan_pairs(c, t, wp, 0, surf));
  }
  if (wp.npair == 0) return RBC3D_OK;
  EvalArgs a = eval_args(c, t, wp);
  k_wall_eval<false><<<(int)(((size_t)wp.npair * 16 + 127) / 128), 128, 0, c->stream>>>(a);
  KERNEL_CHECK();
  k_wall_reduce<<<(t.n + 255) / 256, 256, 0, c->stream>>>(t.n, wp.off.p, wp.npair, wp.dv.p, c1, t.acc.p);
  KERNEL_CHECK();
  c->launches += 2;
  return RBC3D_OK;
}

// in-range (target, element) sets of AddIntOnWalls: count, order-independent checksum, Duffy count per target
__global__ void k_wall_signature(int n, const int *__restrict__ off, const int *__restrict__ ele,
                                 const int *__restrict__ kind, int *__restrict__ count,
                                 unsigned long long *__restrict__ sig, int *__restrict__ nduffy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long s = 0;
  int nd = 0;
  for (int p = off[i]; p < off[i + 1]; p++) {
    unsigned long long z = (unsigned long long)ele[p] + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    s += z ^ (z >> 31);
    nd += kind[p];
  }
  count[i] = off[i + 1] - off[i];
  sig[i] = s;
  nduffy[i] = nd;
}

int walls_signature(rbc3d_ctx *c, TargetList &t, int self_skip, int *count, unsigned long long *sig, int *nduffy) {
  Walls &W = c->walls;
  if (!W.geom_set) return RBC3D_ESTATE;
  WallPairs wp;
  const int *surf = nullptr;
  if (self_skip) {
    RBC_TRY(wall_surf(c, t, &surf));
  } else {
    // no exclusion: compare against a surface id no element has
    TargetList raw;
    raw.kind = RBC3D_TL_RAW;
    raw.n = t.n;
    RBC_TRY(wall_surf(c, raw, &surf));
  }
  RBC_TRY(scan_pairs(c, t, wp, 0, surf));
  const int n = t.n;
  dbuf<int> dc, dn;
  dbuf<unsigned long long> ds;
  RBC_TRY(dc.resize(n > 0 ? n : 1));
  RBC_TRY(dn.resize(n > 0 ? n : 1));
  RBC_TRY(ds.resize(n > 0 ? n : 1));
  if (n > 0) {
    k_wall_signature<<<(n + 255) / 256, 256, 0, c->stream>>>(n, wp.off.p, wp.ele.p, wp.kind.p, dc.p, ds.p, dn.p);
    KERNEL_CHECK();
    CUDA_TRY(cudaMemcpyAsync(count, dc.p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(nduffy, dn.p, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(sig, ds.p, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost, c->stream));
  }
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  dc.release(), dn.release(), ds.release();
  wp.release();
  return RBC3D_OK;
}

// ---------------------------------------------------------------------------------------------------------
// batched scalar entry points (Tri_Int_Regular / Tri_Int_Duffy / MinDistToTri are public in the reference)
__global__ void k_min_dist_batch(int n, const double *__restrict__ xtar, const double *__restrict__ xtri,
                                 double *__restrict__ dist, double *__restrict__ s0, double *__restrict__ t0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double xt[3] = {xtar[i], xtar[(size_t)n + i], xtar[2 * (size_t)n + i]};
  double x[3][3];
#pragma unroll
  for (int l = 0; l < 3; l++)
#pragma unroll
    for (int d = 0; d < 3; d++) x[l][d] = xtri[(size_t)i * 9 + 3 * l + d];
  double s, t;
  dist[i] = min_dist_to_tri(xt, x, s, t);
  s0[i] = s;
  t0[i] = t;
}

int walls_min_dist_batch(rbc3d_ctx *c, int n, const double *xtar, const double *xtri, double *dist, double *s0,
                         double *t0) {
  if (n <= 0) return RBC3D_OK;
  dbuf<double> dx, dt, dd, ds, dtt;
  RBC_TRY(dx.resize(3 * (size_t)n));
  RBC_TRY(dt.resize(9 * (size_t)n));
  RBC_TRY(dd.resize(n));
  RBC_TRY(ds.resize(n));
  RBC_TRY(dtt.resize(n));
  CUDA_TRY(cudaMemcpyAsync(dx.p, xtar, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(dt.p, xtri, sizeof(double) * 9 * n, cudaMemcpyHostToDevice, c->stream));
  k_min_dist_batch<<<(n + 127) / 128, 128, 0, c->stream>>>(n, dx.p, dt.p, dd.p, ds.p, dtt.p);
  KERNEL_CHECK();
  c->launches++;
  CUDA_TRY(cudaMemcpyAsync(dist, dd.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  if (s0) CUDA_TRY(cudaMemcpyAsync(s0, ds.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  if (t0) CUDA_TRY(cudaMemcpyAsync(t0, dtt.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  dx.release(), dt.release(), dd.release(), ds.release(), dtt.release();
  return RBC3D_OK;
}

// n independent (triangle, traction, target) triples: the triangles become a throw-away wall of n elements
int walls_tri_int_batch(rbc3d_ctx *c, int n, const double *xtri, const double *ftri, const double *xtar,
                        const double *s0, const double *t0, double *rhs, double *lhs) {
  if (n <= 0) return RBC3D_OK;
  RBC_TRY(upload_quad(c));
  std::vector<double> hx(9 * (size_t)n), hf(9 * (size_t)n);
  std::vector<int> he(3 * (size_t)n), hw(n, 0), hk(n, s0 ? 1 : 0), hp(n);
  const size_t NV = 3 * (size_t)n;
  for (int i = 0; i < n; i++) {
    hp[i] = i;
    for (int l = 0; l < 3; l++) {
      he[(size_t)l * n + i] = 3 * i + l;
      for (int d = 0; d < 3; d++) {
        hx[(size_t)d * NV + 3 * i + l] = xtri[(size_t)i * 9 + 3 * l + d];
        hf[(size_t)d * NV + 3 * i + l] = ftri ? ftri[(size_t)i * 9 + 3 * l + d] : 0.0;
      }
    }
  }
  dbuf<double> dx, df, dtar, ds, dt, dout;
  dbuf<int> de, dw, dk, dp;
  RBC_TRY(dx.resize(hx.size()));
  RBC_TRY(df.resize(hf.size()));
  RBC_TRY(dtar.resize(3 * (size_t)n));
  RBC_TRY(ds.resize(n));
  RBC_TRY(dt.resize(n));
  RBC_TRY(dout.resize((size_t)n * (lhs ? 27 : 3)));
  RBC_TRY(de.resize(he.size()));
  RBC_TRY(dw.resize(n));
  RBC_TRY(dk.resize(n));
  RBC_TRY(dp.resize(n));
  CUDA_TRY(cudaMemcpy(dx.p, hx.data(), sizeof(double) * hx.size(), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(df.p, hf.data(), sizeof(double) * hf.size(), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dtar.p, xtar, sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemset(ds.p, 0, sizeof(double) * n));
  CUDA_TRY(cudaMemset(dt.p, 0, sizeof(double) * n));
  if (s0) CUDA_TRY(cudaMemcpy(ds.p, s0, sizeof(double) * n, cudaMemcpyHostToDevice));
  if (t0) CUDA_TRY(cudaMemcpy(dt.p, t0, sizeof(double) * n, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(de.p, he.data(), sizeof(int) * he.size(), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dw.p, hw.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dk.p, hk.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dp.p, hp.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
  EvalArgs a;
  a.prm = c->prm;
  // the scalar routines do not translate the triangle: a huge box makes nint(...) = 0
  for (int d = 0; d < 3; d++) a.prm.iLb[d] = 0.0;
  a.g.NV = (int)NV;
  a.g.NE = n;
  a.g.x = dx.p;
  a.g.e2v = de.p;
  a.g.ewall = dw.p;
  a.g.epsDist = nullptr;
  a.npair = n;
  a.ptarget = dp.p;
  a.ele = dp.p;
  a.kind = dk.p;
  a.s0 = ds.p;
  a.t0 = dt.p;
  a.tx = dtar.p;
  a.n = n;
  a.f = df.p;
  a.tab_sl = c->tab_sl.p;
  a.dv = dout.p;
  a.lhs = dout.p;
  const int grid = (int)(((size_t)n * 16 + 127) / 128);
  if (lhs)
    k_wall_eval<true><<<grid, 128, 0, c->stream>>>(a);
  else
    k_wall_eval<false><<<grid, 128, 0, c->stream>>>(a);
  KERNEL_CHECK();
  c->launches++;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (lhs) {
    CUDA_TRY(cudaMemcpy(lhs, dout.p, sizeof(double) * 27 * n, cudaMemcpyDeviceToHost));
  } else {
    std::vector<double> tmp(3 * (size_t)n);
    CUDA_TRY(cudaMemcpy(tmp.data(), dout.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++)
      for (int k = 0; k < 3; k++) rhs[(size_t)i * 3 + k] = tmp[(size_t)k * n + i];
  }
  dx.release(), df.release(), dtar.release(), ds.release(), dt.release(), dout.release();
  de.release(), dw.release(), dk.release(), dp.release();
  return RBC3D_OK;
}

void walls_release(rbc3d_ctx *c) {
  Walls &W = c->walls;
  for (dbuf<double> *b : {&W.x, &W.f, &W.area, &W.epsDist, &W.xc, &W.ft, &W.val}) b->release();
  for (dbuf<int> *b : {&W.e2v, &W.ewall, &W.vwall, &W.src_own, &W.rowptr, &W.col, &W.minus1}) b->release();
  for (CellList *l : {&W.cl, &W.pl}) {
    l->cid.release(), l->order.release(), l->start.release(), l->keys_tmp.release(), l->vals_tmp.release();
    l->cub_tmp.release();
    l->w.release();
  }
  for (int k = 0; k < 3; k++) c->tl[k].wp.release();
}

// ---------------------------------------------------------------------------------------------------------
// Closest_Neighbor_Wall (ModRepulsion.F90:556-613) on the wall cell list: one thread per query point,
// every element whose centroid lies in the 27 neighbouring list cells and that is not on the query's own surface; the
// query point is translated next to the element (:593-595), MinDistToTri gives distance and closest point (:597).
struct ClosestWArgs {
  Params prm;
  WallGeom g;
  int n, id0;                 // id0 = surface id of the first wall (walls(1)%id = nrbc + 1)
  const double *qx;
  const int *surf;
  const int *wstart, *worder;
  double *dist, *x0;
};

__global__ void __launch_bounds__(128) k_closest_wall(ClosestWArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const double xi[3] = {a.qx[i], a.qx[(size_t)a.n + i], a.qx[2 * (size_t)a.n + i]};
  const int *Nc = a.prm.Nc;
  const int i1 = cell_coord(xi[0], a.prm.iLbNc[0], Nc[0]), i2 = cell_coord(xi[1], a.prm.iLbNc[1], Nc[1]),
            i3 = cell_coord(xi[2], a.prm.iLbNc[2], Nc[2]);
  const int sid = a.surf[i];
  double best = INFINITY, bx[3] = {0, 0, 0};
  int be = 0x7fffffff;
  for (int d3 = -1; d3 <= 1; d3++)
    for (int d2 = -1; d2 <= 1; d2++)
      for (int d1 = -1; d1 <= 1; d1++) {
        const int cj = imodulo(i1 + d1, Nc[0]) + Nc[0] * (imodulo(i2 + d2, Nc[1]) + Nc[1] * imodulo(i3 + d3, Nc[2]));
        for (int k = a.wstart[cj]; k < a.wstart[cj + 1]; k++) {
          const int e = a.worder[k];
          if (a.id0 + a.g.ewall[e] == sid) continue;  // :582
          double x[3][3], xt[3], s0, t0;
#pragma unroll
          for (int l = 0; l < 3; l++) {
            const int iv = a.g.e2v[(size_t)l * a.g.NE + e];
#pragma unroll
            for (int d = 0; d < 3; d++) x[l][d] = a.g.x[(size_t)d * a.g.NV + iv];
          }
#pragma unroll
          for (int d = 0; d < 3; d++) {
            double xx = sub_(xi[d], x[0][d]);
            xx = sub_(xx, mul_(round(mul_(xx, a.prm.iLb[d])), a.prm.Lb[d]));
            xt[d] = add_(x[0][d], xx);
          }
          const double rr = min_dist_to_tri(xt, x, s0, t0);
          if (rr < best || (rr == best && e < be)) {
            best = rr, be = e;
#pragma unroll
            for (int d = 0; d < 3; d++) bx[d] = (1.0 - s0 - t0) * x[0][d] + s0 * x[1][d] + t0 * x[2][d];  // ModIntOnWalls.F90:575
          }
        }
      }
  a.dist[i] = best;
#pragma unroll
  for (int d = 0; d < 3; d++) a.x0[(size_t)d * a.n + i] = bx[d];
}

int closest_walls(rbc3d_ctx *c, int n, const double *qx, const int *surf, double *dist, double *x0) {
  Walls &W = c->walls;
  if (n == 0) return RBC3D_OK;
  ClosestWArgs a;
  a.prm = c->prm;
  a.g.NV = W.NV, a.g.NE = W.NE, a.g.x = W.x.p, a.g.e2v = W.e2v.p, a.g.ewall = W.ewall.p, a.g.epsDist = W.epsDist.p;
  a.n = n, a.id0 = c->cells.ncell + 1;
  a.qx = qx, a.surf = surf, a.wstart = W.cl.start.p, a.worder = W.cl.order.p;
  a.dist = dist, a.x0 = x0;
  k_closest_wall<<<(n + 127) / 128, 128, 0, c->stream>>>(a);
  KERNEL_CHECK();
  c->launches++;
  return RBC3D_OK;
}

}  // namespace rbc3d
