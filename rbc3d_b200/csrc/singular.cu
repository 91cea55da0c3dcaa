// singular.cu -- singular self-cell correction of the real-space sum: RBC_SingInt (ModRbcSingInt.F90:29-90),
// called from AddIntOnRbcs for every on-surface target (ModIntOnRbcs.F90:114-121).
//
// For a target (cell, ilat0, ilon0) the masked-out part of the pair sum is replaced by quadrature on a polar
// patch of nrad x nazm points whose reference-sphere coordinates thG/phiG are shared by all cells; positions,
// normals and densities at the patch points come from the cell's bicubic splines (Spline_Interp,
// ModSpline.F90:150-191).
//
// Two paths:
//  * k_singular (direct): one warp per target, everything evaluated from the ABI-layout splines.  Used for the
//    single-layer operator (once per time step) and whenever the cached path is not available.
//  * cached double-layer path (the GMRES matvec, c1 = 0): everything that does not depend on the density is
//    hoisted out of the matvec --
//      - set_mesh time, cell independent: targets grouped in tiles of 4 lat x 2 lon mesh points; per tile the
//        bounding window of spline nodes its 8 patches touch; per patch point the window-relative node index and
//        the fractional coordinates (s, t) of the bicubic cell;
//      - geometry time, per cell: xx = x(patch point) - x(target) and s = w * EwaldCoeff_DL(|xx|) * (xx . a3) per
//        patch point, 32 B each, streamed from HBM by the matvec (k_sing_cache_build);
//      - per matvec: spline(g detJ) is re-laid out node-interleaved ([cell][half][phi][theta][6 doubles], 48 B
//        node halves), one CTA per (cell, tile row) stages the band of the spline the row's patches touch and walks
//        over the row's tiles.  The 8 x 288 patch points of a tile are processed SORTED BY SPLINE CELL (a cell-independent permutation, so
//        is the cache layout): the 32 lanes of a warp then read only 3-6 distinct spline cells per instruction
//        and the LDS.128 node gathers are served as multicasts (5-12 patch points share a spline cell) instead of
//        32 distinct 48-byte records.  Every point's contribution w EA (xx.a3)(xx.g) xx is written to a
//        target-major shared buffer, which one warp per target then sums in the patch order of the reference
//        (fixed order: deterministic).
#include <algorithm>
#include <cstdlib>

#include "device_math.cuh"
#include "rbc3d_internal.h"

namespace rbc3d {

constexpr int SING_WARPS = 8;
constexpr int SG_TLAT = 4, SG_TLON = 2, SG_T = SG_TLAT * SG_TLON;  // targets per tile = warps per CTA
constexpr size_t SG_SMEM_MAX = 226 * 1024;                          // one CTA per SM

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

// ---------------------------------------------------------------------------------------------------------
// cell-independent tables (host, once per rbc3d_cells_set_mesh)

// smallest cyclic interval [lo, lo+len) of Z_mod that contains every marked index
static void cyclic_cover(const std::vector<char> &used, int mod, int &lo, int &len) {
  int nused = 0;
  for (int i = 0; i < mod; i++) nused += used[i] ? 1 : 0;
  if (nused == 0) {
    lo = 0;
    len = 1;
    return;
  }
  // largest run of unused indices (cyclic) -> the cover is its complement
  int best_len = -1, best_start = 0;
  for (int s = 0; s < mod; s++) {
    if (used[s]) continue;
    if (!used[(s + mod - 1) % mod] || nused == 0) continue;  // only starts of gaps
    int l = 0;
    while (l < mod && !used[(s + l) % mod]) l++;
    if (l > best_len) {
      best_len = l;
      best_start = s;
    }
  }
  if (best_len <= 0) {
    lo = 0;
    len = mod;
    return;
  }
  lo = (best_start + best_len) % mod;
  len = mod - best_len;
}

int singular_mesh_prepare(rbc3d_ctx *c, const double *thG, const double *phiG) {
  Cells &C = c->cells;
  C.sg_ok = false;
  C.sg_cache_ok = false;
  const int nlat = C.nlat, nlon = C.nlon, m = 2 * nlat, n = nlon, npatch = C.nrad * C.nazm;
  const int ntl = (nlat + SG_TLAT - 1) / SG_TLAT, ntn = nlon / SG_TLON;
  const int K = (npatch + 31) / 32, NPT = K * SG_T * 32;
  const double hx = RBC_TWO_PI / (double)m, hy = RBC_TWO_PI / (double)n;
  const double ihx = 1.0 / hx, ihy = 1.0 / hy;
  C.sg_ntl = ntl;
  C.sg_ntn = ntn;
  C.sg_ntiles = ntl * ntn;
  C.sg_K = K;
  if (nlon % SG_TLON != 0 || n > 1023 || NPT > 16383) return RBC3D_OK;  // direct kernel only
  // Tables of tile column 0 of every tile row.  The patch of (ilat, ilon) is the patch of (ilat, 0) rotated by
  // phi(ilon) about the polar axis (PolarPatch_Build, ModPolarPatch.F90:99-148: thG does not depend on phi0, phiG
  // = atan2(..) + phi0), so tile column tn uses the same tables with the phi node index advanced by tn*SG_TLON.
  std::vector<int> row_tgt((size_t)ntl * SG_T, -1), row_win((size_t)ntl * 2, 0);
  std::vector<int> pk((size_t)ntl * NPT, 0), pos((size_t)ntl * NPT, 0);
  std::vector<double> st((size_t)ntl * NPT * 2, 0.0);
  std::vector<int> ni1((size_t)SG_T * npatch), nj1((size_t)SG_T * npatch);
  std::vector<double> fs((size_t)SG_T * npatch), ft((size_t)SG_T * npatch);
  std::vector<std::pair<long long, int>> keys(NPT);
  int ni_max = 0;
  for (int tl = 0; tl < ntl; tl++) {
    std::vector<char> ui(m, 0);
    for (int w = 0; w < SG_T; w++) {
      const int ilat = tl * SG_TLAT + (w % SG_TLAT), ilon = w / SG_TLAT;
      if (ilat >= nlat) continue;
      const int p = ilon * nlat + ilat;
      row_tgt[(size_t)tl * SG_T + w] = p;
      for (int q = 0; q < npatch; q++) {
        // same arithmetic as spline_interp (device_math.cuh)
        const double xs = thG[(size_t)p * npatch + q] * ihx, ys = phiG[(size_t)p * npatch + q] * ihy;
        const int i1 = (int)floor(xs), j1 = (int)floor(ys);
        fs[(size_t)w * npatch + q] = xs - (double)i1;
        ft[(size_t)w * npatch + q] = ys - (double)j1;
        const int i1m = ((i1 % m) + m) % m, j1m = ((j1 % n) + n) % n;
        ni1[(size_t)w * npatch + q] = i1m;
        nj1[(size_t)w * npatch + q] = j1m;
        ui[i1m] = ui[(i1m + 1) % m] = 1;
      }
    }
    int ilo, ni;
    cyclic_cover(ui, m, ilo, ni);
    if (ni == m) ni = m + 1;  // a full circle needs the first row once more so that node+1 stays inside
    row_win[(size_t)tl * 2 + 0] = ilo;
    row_win[(size_t)tl * 2 + 1] = ni;
    ni_max = std::max(ni_max, ni);
    // sorted order of the tile's patch points: by spline cell (phi column, theta row), ties in target-major order
    for (int k = 0; k < K; k++)
      for (int w = 0; w < SG_T; w++)
        for (int lane = 0; lane < 32; lane++) {
          const int slot = (k * SG_T + w) * 32 + lane, q = k * 32 + lane;
          const bool valid = row_tgt[(size_t)tl * SG_T + w] >= 0 && q < npatch;
          long long key = 1LL << 40;
          if (valid) {
            const int wi = (ni1[(size_t)w * npatch + q] - ilo + m) % m;
            key = (long long)nj1[(size_t)w * npatch + q] * ni + wi;
          }
          keys[slot] = {key * (long long)NPT + (w * K * 32 + q), slot};
        }
    std::sort(keys.begin(), keys.end());
    for (int sp = 0; sp < NPT; sp++) {
      const int slot = keys[sp].second;
      const int lane = slot % 32, w = (slot / 32) % SG_T, k = slot / (32 * SG_T), q = k * 32 + lane;
      const int dest = w * K * 32 + q;
      const bool valid = row_tgt[(size_t)tl * SG_T + w] >= 0 && q < npatch;
      int wi = 0, j0 = 0;
      if (valid) {
        wi = (ni1[(size_t)w * npatch + q] - ilo + m) % m;
        j0 = nj1[(size_t)w * npatch + q];
        st[2 * ((size_t)tl * NPT + sp)] = fs[(size_t)w * npatch + q];
        st[2 * ((size_t)tl * NPT + sp) + 1] = ft[(size_t)w * npatch + q];
      }
      pk[(size_t)tl * NPT + sp] = wi | (j0 << 8) | (dest << 18);
      pos[(size_t)tl * NPT + slot] = sp;
    }
  }
  C.sg_ni_max = ni_max;
  // shared memory: band of the spline (2 halves x (n+1) phi columns x ni theta rows x 48 B) + contribution buffer
  const size_t smem = (size_t)12 * ni_max * (n + 1) * sizeof(double) + (size_t)3 * NPT * sizeof(double);
  if (smem > SG_SMEM_MAX || ni_max > 255) return RBC3D_OK;  // direct kernel only
  C.sg_smem = smem;
  RBC_TRY(C.sg_tile_tgt.resize(row_tgt.size()));
  RBC_TRY(C.sg_tile_win.resize(row_win.size()));
  RBC_TRY(C.sg_idx.resize(pk.size()));
  RBC_TRY(C.sg_pos.resize(pos.size()));
  RBC_TRY(C.sg_st.resize(st.size()));
  CUDA_TRY(cudaMemcpyAsync(C.sg_tile_tgt.p, row_tgt.data(), sizeof(int) * row_tgt.size(), cudaMemcpyHostToDevice,
                           c->stream));
  CUDA_TRY(cudaMemcpyAsync(C.sg_tile_win.p, row_win.data(), sizeof(int) * row_win.size(), cudaMemcpyHostToDevice,
                           c->stream));
  CUDA_TRY(cudaMemcpyAsync(C.sg_idx.p, pk.data(), sizeof(int) * pk.size(), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(C.sg_pos.p, pos.data(), sizeof(int) * pos.size(), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(C.sg_st.p, st.data(), sizeof(double) * st.size(), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  C.sg_ok = true;
  return RBC3D_OK;
}

// ---------------------------------------------------------------------------------------------------------
// geometry time: density-independent factors of every patch point (one warp per target), written in the sorted
// order of the target's tile
struct CacheArgs {
  Params prm;
  int ncell, npc, nlat, nlon, npatch, nrad, ntn, K;
  const double *th, *phi, *thG, *phiG, *pw;
  const double *spx, *spa3;
  const int *row_tgt;
  const int *pos;  // [tile row][K][T][32] -> position in the tile's sorted order
  const double *tab_dl;
  double4 *cache;  // [cell][tile][sorted position]
};

__global__ void __launch_bounds__(SG_T * 32) k_sing_cache_build(CacheArgs a) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, cell = blockIdx.y;
  const int tl = tile / a.ntn, tn = tile - tl * a.ntn;
  const int pt0 = a.row_tgt[tl * SG_T + w];
  const int NPT = a.K * SG_T * 32;
  double4 *out = a.cache + ((size_t)cell * gridDim.x + tile) * NPT;
  const int *pos = a.pos + (size_t)tl * NPT;
  if (pt0 < 0) {
    for (int k = 0; k < a.K; k++) out[pos[(k * SG_T + w) * 32 + lane]] = make_double4(0, 0, 0, 0);
    return;
  }
  const int pt = pt0 + tn * SG_TLON * a.nlat;
  const int ilon0 = pt / a.nlat, ilat0 = pt - ilon0 * a.nlat;
  const int m = 2 * a.nlat, n = a.nlon;
  const size_t sp3 = (size_t)12 * m * n;
  const double *spx = a.spx + sp3 * cell;
  const double *spa3 = a.spa3 + sp3 * cell;
  double xi[3];
  spline_interp<3>(spx, m, n, a.th[ilat0], a.phi[ilon0], xi);
  const size_t off = (size_t)pt * a.npatch;
  for (int k = 0; k < a.K; k++) {
    const int q = k * 32 + lane;
    double4 r = make_double4(0, 0, 0, 0);
    if (q < a.npatch) {
      const double th_j = __ldg(a.thG + off + q), phi_j = __ldg(a.phiG + off + q);
      const double wq = __ldg(a.pw + (q % a.nrad));
      double xj[3], nj[3];
      spline_interp<3>(spx, m, n, th_j, phi_j, xj);
      const double xx = xj[0] - xi[0], yy = xj[1] - xi[1], zz = xj[2] - xi[2];
      const double rr = sqrt(xx * xx + yy * yy + zz * zz);
      if (rr < a.prm.rc) {  // ModRbcSingInt.F90:69
        spline_interp<3>(spa3, m, n, th_j, phi_j, nj);
        const double EA = ewald_dl(a.tab_dl, a.prm, rr);
        r = make_double4(xx, yy, zz, EA * wq * (xx * nj[0] + yy * nj[1] + zz * nj[2]));
      }
    }
    out[pos[(k * SG_T + w) * 32 + lane]] = r;
  }
}

// spline re-layout: ABI [cell][4 (u,u1,u2,u12)][3][n][m] -> [cell][half][n][m][6], half 0 = (u[0..2], u1[0..2]),
// half 1 = (u2[0..2], u12[0..2])
__global__ void __launch_bounds__(256) k_spline_interleave(int ncell, int plane, const double *__restrict__ sp,
                                                           double *__restrict__ out) {
  const size_t total = (size_t)ncell * plane;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t cell = e / plane, node = e - cell * plane;
    const double *src = sp + cell * 12 * (size_t)plane + node;
    double v[12];
#pragma unroll
    for (int a = 0; a < 12; a++) v[a] = src[(size_t)a * plane];  // a = arr*3 + var
    double2 *dA = reinterpret_cast<double2 *>(out + ((cell * 2 + 0) * plane + node) * 6);
    double2 *dB = reinterpret_cast<double2 *>(out + ((cell * 2 + 1) * plane + node) * 6);
    dA[0] = make_double2(v[0], v[1]);
    dA[1] = make_double2(v[2], v[3]);
    dA[2] = make_double2(v[4], v[5]);
    dB[0] = make_double2(v[6], v[7]);
    dB[1] = make_double2(v[8], v[9]);
    dB[2] = make_double2(v[10], v[11]);
  }
}

struct BandArgs {
  int ncell, npc, nlat, nlon, ntl, ntn, Np, K;
  const int *row_tgt;      // [tile row][T]: mesh point (ilon*nlat + ilat) of the targets of tile column 0, -1 = none
  const int *row_win;      // [tile row][2]: first theta row of the band, number of rows
  const int *pk;           // [tile row][sorted position]: theta row | phi column << 8 | target-major slot << 18
  const double2 *st;       // [tile row][sorted position]: fractional coordinates in the spline cell
  const double *spGi;      // [cell][2][n][m][6]
  const double4 *cache;    // [cell][tile][sorted position]
  const double *Bcell;
  const int *active;
  const int *cell_active;  // per cell: any active target
  double c2;
  double *acc;
};

// streaming (evict-first) 32-byte load: the cache is read exactly once per matvec
__device__ __forceinline__ double4 ld_stream4(const double4 *p) {
  const double2 lo = __ldcs(reinterpret_cast<const double2 *>(p));
  const double2 hi = __ldcs(reinterpret_cast<const double2 *>(p) + 1);
  return make_double4(lo.x, lo.y, hi.x, hi.y);
}

// one bicubic evaluation of 3 variables from the staged window: node halves A = (u, u1), B = (u2, u12)
__device__ __forceinline__ void interp_window(const double *__restrict__ sA, const double *__restrict__ sB, int a11,
                                              int ni, const double cx[4], const double cy[4], double g[3]) {
  const int a21 = a11 + 1, a12 = a11 + ni, a22 = a12 + 1;
  const double2 *A11 = reinterpret_cast<const double2 *>(sA + 6 * a11), *A12 = reinterpret_cast<const double2 *>(sA + 6 * a12);
  const double2 *A21 = reinterpret_cast<const double2 *>(sA + 6 * a21), *A22 = reinterpret_cast<const double2 *>(sA + 6 * a22);
  const double2 *B11 = reinterpret_cast<const double2 *>(sB + 6 * a11), *B12 = reinterpret_cast<const double2 *>(sB + 6 * a12);
  const double2 *B21 = reinterpret_cast<const double2 *>(sB + 6 * a21), *B22 = reinterpret_cast<const double2 *>(sB + 6 * a22);
  // per node: (u0,u1,u2, d0,d1,d2) with d = d/dtheta (half A) and (e0,e1,e2, f0,f1,f2) = d/dphi, d2/dthdphi (half B)
  double n11[12], n12[12], n21[12], n22[12];
#pragma unroll
  for (int h = 0; h < 3; h++) {
    double2 v;
    v = A11[h]; n11[2 * h] = v.x; n11[2 * h + 1] = v.y;
    v = B11[h]; n11[6 + 2 * h] = v.x; n11[7 + 2 * h] = v.y;
    v = A12[h]; n12[2 * h] = v.x; n12[2 * h + 1] = v.y;
    v = B12[h]; n12[6 + 2 * h] = v.x; n12[7 + 2 * h] = v.y;
    v = A21[h]; n21[2 * h] = v.x; n21[2 * h + 1] = v.y;
    v = B21[h]; n21[6 + 2 * h] = v.x; n21[7 + 2 * h] = v.y;
    v = A22[h]; n22[2 * h] = v.x; n22[2 * h + 1] = v.y;
    v = B22[h]; n22[6 + 2 * h] = v.x; n22[7 + 2 * h] = v.y;
  }
#pragma unroll
  for (int l = 0; l < 3; l++) {
    // U = n[l], U1 = n[3+l], U2 = n[6+l], U12 = n[9+l]   (same association as spline_interp)
    const double r0 = n11[l] * cy[0] + n12[l] * cy[1] + n11[6 + l] * cy[2] + n12[6 + l] * cy[3];
    const double r1 = n21[l] * cy[0] + n22[l] * cy[1] + n21[6 + l] * cy[2] + n22[6 + l] * cy[3];
    const double r2 = n11[3 + l] * cy[0] + n12[3 + l] * cy[1] + n11[9 + l] * cy[2] + n12[9 + l] * cy[3];
    const double r3 = n21[3 + l] * cy[0] + n22[3 + l] * cy[1] + n21[9 + l] * cy[2] + n22[9 + l] * cy[3];
    g[l] = cx[0] * r0 + cx[1] * r1 + cx[2] * r2 + cx[3] * r3;
  }
}

// One CTA per (cell, tile row): the band of spline(g detJ) that the row's patches touch (all phi columns, ni theta
// rows) is staged once and serves the ntn tiles of the row; the per-point tables live in registers for the whole
// CTA (tile column tn only advances the phi column by tn*SG_TLON); the geometry cache of the NEXT tile is already in
// flight (one register set, refilled as it is consumed) while the current tile is evaluated.
template <int KT>  // patch points per thread (0: generic, tables and cache re-read per tile without the register set)
__global__ void __launch_bounds__(SG_T * 32, 1) k_sing_band(BandArgs a) {
  extern __shared__ double smem[];
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int cell = blockIdx.x / a.ntl, tl = blockIdx.x - cell * a.ntl;
  if (!a.cell_active[cell]) return;  // block-uniform
  const int ilo = a.row_win[tl * 2 + 0], ni = a.row_win[tl * 2 + 1];
  const int m = 2 * a.nlat, n = a.nlon, plane = m * n;
  const int K = a.K, NPT = K * SG_T * 32;
  double *sA = smem, *sB = smem + (size_t)6 * ni * (n + 1);
  double *sC = sB + (size_t)6 * ni * (n + 1);  // [3][NPT] contributions, target-major
  const double hx = RBC_TWO_PI / (double)m, hy = RBC_TWO_PI / (double)n;
  constexpr int KR = KT > 0 ? KT : 1;
  int r_pk[KR];
  double2 r_st[KR];
  double4 r_c[KR];
  const size_t ebase = (size_t)tl * NPT + tid;
  const double4 *cg = a.cache + ((size_t)cell * a.ntl * a.ntn + (size_t)tl * a.ntn) * NPT + tid;
  if (KT > 0) {
#pragma unroll
    for (int k = 0; k < KR; k++) {
      r_c[k] = ld_stream4(cg + (size_t)k * SG_T * 32);
      r_pk[k] = a.pk[ebase + (size_t)k * SG_T * 32];
      r_st[k] = a.st[ebase + (size_t)k * SG_T * 32];
    }
  }
  // stage the band: rows = (half, phi column 0..n with column n = column 0), each a cyclic run of ni nodes x 48 B
  for (int row = w; row < 2 * (n + 1); row += SG_T) {
    const int h = row / (n + 1), wj = row - h * (n + 1);
    const int j = wj == n ? 0 : wj;
    const double2 *src = reinterpret_cast<const double2 *>(a.spGi + (((size_t)cell * 2 + h) * plane + (size_t)j * m) * 6);
    double2 *dst = reinterpret_cast<double2 *>((h ? sB : sA) + (size_t)6 * wj * ni);
    for (int u = lane; u < 3 * ni; u += 32) {
      const int wi = u / 3, part = u - 3 * wi;
      int i = ilo + wi;
      if (i >= m) i -= m;
      dst[u] = __ldg(src + 3 * i + part);
    }
  }
  __syncthreads();
  const int pt0 = a.row_tgt[tl * SG_T + w];
  const double c2m = a.c2 * a.Bcell[cell];  // c2Mod, ModIntOnRbcs.F90:116
  for (int tn = 0; tn < a.ntn; tn++) {
    const int jshift = tn * SG_TLON;
    auto point = [&](int pk, double2 stv, double4 c4) {
      int j = ((pk >> 8) & 1023) + jshift;
      if (j >= n) j -= n;
      const int a11 = j * ni + (pk & 255), dest = pk >> 18;
      const double s = stv.x, t = stv.y;
      const double cx[4] = {1.0 + s * s * (-3.0 + 2.0 * s), s * s * (3.0 - 2.0 * s), hx * s * (1.0 + s * (-2.0 + s)),
                            hx * s * s * (-1.0 + s)};
      const double cy[4] = {1.0 + t * t * (-3.0 + 2.0 * t), t * t * (3.0 - 2.0 * t), hy * t * (1.0 + t * (-2.0 + t)),
                            hy * t * t * (-1.0 + t)};
      double g[3];
      interp_window(sA, sB, a11, ni, cx, cy, g);
      const double qd = c4.w * (c4.x * g[0] + c4.y * g[1] + c4.z * g[2]);
      sC[dest] = qd * c4.x;
      sC[NPT + dest] = qd * c4.y;
      sC[2 * NPT + dest] = qd * c4.z;
    };
    const double4 *cgn = cg + (size_t)(tn + 1) * NPT;  // next tile of the row
    const bool more = tn + 1 < a.ntn;
    if (KT > 0) {
#pragma unroll
      for (int k = 0; k < KR; k++) {
        const double4 c4 = r_c[k];
        if (more) r_c[k] = ld_stream4(cgn + (size_t)k * SG_T * 32);
        point(r_pk[k], r_st[k], c4);
      }
    } else {
      const double4 *cgc = cg + (size_t)tn * NPT;
      double4 c_next = ld_stream4(cgc);
      for (int k = 0; k < K; k++) {
        const double4 c4 = c_next;
        if (k + 1 < K) c_next = ld_stream4(cgc + (size_t)(k + 1) * SG_T * 32);
        point(a.pk[ebase + (size_t)k * SG_T * 32], a.st[ebase + (size_t)k * SG_T * 32], c4);
      }
    }
    __syncthreads();
    // one warp per target: sum its patch in the reference's order (lane = point mod 32, then the shuffle tree)
    if (pt0 >= 0) {
      double dvx = 0, dvy = 0, dvz = 0;
      const double *sc = sC + w * K * 32 + lane;
      for (int k = 0; k < K; k++) {
        dvx += sc[k * 32];
        dvy += sc[NPT + k * 32];
        dvz += sc[2 * NPT + k * 32];
      }
      dvx = warp_sum(dvx);
      dvy = warp_sum(dvy);
      dvz = warp_sum(dvz);
      if (lane == 0) {
        const int ti = cell * a.npc + pt0 + tn * SG_TLON * a.nlat;
        if (a.active[ti]) {
          a.acc[ti] += c2m * dvx;
          a.acc[(size_t)a.Np + ti] += c2m * dvy;
          a.acc[2 * (size_t)a.Np + ti] += c2m * dvz;
        }
      }
    }
    __syncthreads();
  }
}

__global__ void k_cell_active(int Np, int npc, const int *__restrict__ active, int *__restrict__ cell_active) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Np && active[i]) cell_active[i / npc] = 1;
}

// per cell: does it own any active target (z-slab / cell ownership of multi-GPU runs, SetActiveFlag)
int cells_active_flags(rbc3d_ctx *c) {
  Cells &C = c->cells;
  TargetList &t = c->tl[RBC3D_TL_CELLS];
  RBC_TRY(C.sg_cell_active.resize(C.ncell > 0 ? C.ncell : 1));
  if (C.Np == 0) return RBC3D_OK;
  CUDA_TRY(cudaMemsetAsync(C.sg_cell_active.p, 0, sizeof(int) * C.ncell, c->stream));
  k_cell_active<<<(C.Np + 255) / 256, 256, 0, c->stream>>>(C.Np, C.npc, t.active.p, C.sg_cell_active.p);
  KERNEL_CHECK();
  c->launches++;
  return RBC3D_OK;
}

// geometry time (SourceList_UpdateCoord): density-independent cache of the double-layer patch integrand
int singular_prepare(rbc3d_ctx *c) {
  Cells &C = c->cells;
  C.sg_cache_ok = false;
  if (!C.sg_ok || C.Np == 0 || c->sing_cache_mode == 0) return RBC3D_OK;
  const size_t per_cell = (size_t)C.sg_ntiles * C.sg_K * SG_T * 32;
  const size_t need = per_cell * C.ncell * sizeof(double4);
  if (C.sg_cache.n < per_cell * C.ncell) {
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    // leave room for the interleaved density spline and the rest of the working set
    const size_t reserve = (size_t)C.ncell * 12 * 2 * C.nlat * C.nlon * 8 * 2 + ((size_t)2 << 30);
    if (need + reserve > free_b + C.sg_cache.n * sizeof(double4)) return RBC3D_OK;  // direct kernel
    if (C.sg_cache.resize(per_cell * C.ncell) != RBC3D_OK) return RBC3D_OK;
  }
  CacheArgs a;
  a.prm = c->prm;
  a.ncell = C.ncell;
  a.npc = C.npc;
  a.nlat = C.nlat;
  a.nlon = C.nlon;
  a.npatch = C.nrad * C.nazm;
  a.nrad = C.nrad;
  a.ntn = C.sg_ntn;
  a.K = C.sg_K;
  a.th = C.th.p;
  a.phi = C.phi.p;
  a.thG = C.thG.p;
  a.phiG = C.phiG.p;
  a.pw = C.pw.p;
  a.spx = C.spx.p;
  a.spa3 = C.spa3.p;
  a.row_tgt = C.sg_tile_tgt.p;
  a.pos = C.sg_pos.p;
  a.tab_dl = c->tab_dl.p;
  a.cache = C.sg_cache.p;
  // 65535 limit of gridDim.y: chunk the cells
  for (int c0 = 0; c0 < C.ncell; c0 += 32768) {
    CacheArgs b = a;
    const int nc = std::min(32768, C.ncell - c0);
    b.spx = a.spx + (size_t)12 * 2 * C.nlat * C.nlon * c0;
    b.spa3 = a.spa3 + (size_t)12 * 2 * C.nlat * C.nlon * c0;
    b.cache = a.cache + per_cell * c0;
    k_sing_cache_build<<<dim3(C.sg_ntiles, nc), SG_T * 32, 0, c->stream>>>(b);
    KERNEL_CHECK();
    c->launches++;
  }
  C.sg_cache_ok = true;
  if (C.g_set && !C.spGi_valid) RBC_TRY(singular_density_prepare(c));
  return RBC3D_OK;
}

// density time (SourceList_UpdateDensity): node-interleaved copy of spline(g detJ) for the cached path
int singular_density_prepare(rbc3d_ctx *c) {
  Cells &C = c->cells;
  if (!C.sg_cache_ok || !C.g_set || C.Np == 0) return RBC3D_OK;
  const int plane = 2 * C.nlat * C.nlon;
  RBC_TRY(C.spGi.resize((size_t)C.ncell * 12 * plane));
  k_spline_interleave<<<c->sm_count * 8, 256, 0, c->stream>>>(C.ncell, plane, C.spG.p, C.spGi.p);
  KERNEL_CHECK();
  c->launches++;
  C.spGi_valid = true;
  return RBC3D_OK;
}

static int singular_apply_cached(rbc3d_ctx *c, TargetList &t, double c2) {
  Cells &C = c->cells;
  BandArgs a;
  a.ncell = C.ncell;
  a.npc = C.npc;
  a.nlat = C.nlat;
  a.nlon = C.nlon;
  a.ntl = C.sg_ntl;
  a.ntn = C.sg_ntn;
  a.Np = C.Np;
  a.K = C.sg_K;
  a.row_tgt = C.sg_tile_tgt.p;
  a.row_win = C.sg_tile_win.p;
  a.pk = C.sg_idx.p;
  a.st = reinterpret_cast<const double2 *>(C.sg_st.p);
  a.spGi = C.spGi.p;
  a.cache = C.sg_cache.p;
  a.Bcell = C.B.p;
  a.active = t.active.p;
  a.cell_active = C.sg_cell_active.p;
  a.c2 = c2;
  a.acc = t.acc.p;
  static const bool generic = getenv("RBC3D_SING_GENERIC") != nullptr;
  const int grid = C.ncell * C.sg_ntl;
  const size_t smem = C.sg_smem;
  if (C.sg_K == 9 && !generic) {
    CUDA_TRY(cudaFuncSetAttribute(k_sing_band<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_sing_band<9><<<grid, SG_T * 32, smem, c->stream>>>(a);
  } else if (C.sg_K == 4 && !generic) {
    CUDA_TRY(cudaFuncSetAttribute(k_sing_band<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_sing_band<4><<<grid, SG_T * 32, smem, c->stream>>>(a);
  } else {
    CUDA_TRY(cudaFuncSetAttribute(k_sing_band<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_sing_band<0><<<grid, SG_T * 32, smem, c->stream>>>(a);
  }
  KERNEL_CHECK();
  c->launches++;
  return RBC3D_OK;
}

int singular_apply(rbc3d_ctx *c, TargetList &t, double c1, double c2) {
  if (t.kind != RBC3D_TL_CELLS) return RBC3D_OK;  // only on-surface targets (ModIntOnRbcs.F90:115)
  Cells &C = c->cells;
  if (C.Np == 0) return RBC3D_OK;
  const bool sl = (c1 != 0), dl = (c2 != 0);
  if (!sl && !dl) return RBC3D_OK;
  if (!sl && dl && C.sg_cache_ok && C.spGi_valid) return singular_apply_cached(c, t, c2);
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
