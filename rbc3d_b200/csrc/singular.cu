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
constexpr size_t SG_SMEM_MAX = 227 * 1024;                          // one CTA per SM
constexpr int SG_PC_DEFAULT = 2;  // patch points of one spline cell evaluated per thread from one load of its 4 nodes
// 0 (default): the tile's points sorted by spline cell across its targets + shared contribution buffer; 1: every warp
// owns one target of the tile (register accumulation, no contribution buffer, no barriers in the tile loop).  Measured
// at 512 cells: 7.8 ms vs 8.1 ms (profiles/r01_summary)
static int sg_lr();
static int sg_pitch_host(int ni);
static int sg_per_warp() {
  static const int v = [] {
    const char *e = getenv("RBC3D_SING_PER_WARP");
    return (e ? atoi(e) : 0) || sg_lr() == 2;
  }();
  return v;
}
static int sg_lr() {
  static const int v = [] {
    const char *e = getenv("RBC3D_SING_LR");  // 0: 256 threads, full-register body; 1: 512 threads, tile mode
    return e ? atoi(e) : 2;                   // 2 (default): 512 threads, warp = target, low-register body
  }();
  return v;
}
static int sg_nt() {
  static const int v = [] {
    const char *e = getenv("RBC3D_SING_NT");
    const int q = e ? atoi(e) : SG_T * 32;
    if (sg_lr() == 1 && !sg_per_warp()) return 512;
    return (!sg_per_warp() && (q == 256 || q == 384 || q == 512)) ? q : SG_T * 32;
  }();
  return v;
}
static int sg_pc() {
  static const int v = [] {
    const char *e = getenv("RBC3D_SING_PC");
    const int q = e ? atoi(e) : SG_PC_DEFAULT;
    return (q == 2 || q == 3 || q == 4 || q == 6) ? q : SG_PC_DEFAULT;
  }();
  return v;
}

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

int singular_mesh_prepare(rbc3d_ctx *c, const double *thG, const double *phiG, const double *pw) {
  Cells &C = c->cells;
  C.sg_ok = false;
  C.sg_cache_ok = false;
  const int nlat = C.nlat, nlon = C.nlon, m = 2 * nlat, n = nlon, npatch = C.nrad * C.nazm;
  const int ntl = (nlat + SG_TLAT - 1) / SG_TLAT, ntn = nlon / SG_TLON;
  const int K = (npatch + 31) / 32, NPT = K * SG_T * 32;
  const double hx = RBC_TWO_PI / (double)m, hy = RBC_TWO_PI / (double)n;
  const double ihx = 1.0 / hx, ihy = 1.0 / hy;
  C.sg_npatch_active = 0;
  for (int q = 0; q < npatch; q++) C.sg_npatch_active += (pw[q % C.nrad] != 0.0) ? 1 : 0;
  C.sg_ntl = ntl;
  C.sg_ntn = ntn;
  C.sg_ntiles = ntl * ntn;
  C.sg_K = K;
  if (nlon % SG_TLON != 0 || n > 1023 || NPT > 16383) return RBC3D_OK;  // direct kernel only
  // Tables of tile column 0 of every tile row.  The patch of (ilat, ilon) is the patch of (ilat, 0) rotated by
  // phi(ilon) about the polar axis (PolarPatch_Build, ModPolarPatch.F90:99-148: thG does not depend on phi0, phiG
  // = atan2(..) + phi0), so tile column tn uses the same tables with the phi node index advanced by tn*SG_TLON.
  const int pc_max = sg_pc();
  const bool per_warp = sg_per_warp() != 0;
  std::vector<int> row_tgt((size_t)ntl * SG_T, -1), row_win((size_t)ntl * 2, 0), row_rounds(ntl, 0);
  std::vector<int> pt_dest((size_t)ntl * NPT, 0), pos((size_t)ntl * NPT, -1);
  std::vector<double> st((size_t)ntl * NPT * 2, 0.0);
  std::vector<std::vector<int2>> chunks(ntl);
  std::vector<int> ni1((size_t)SG_T * npatch), nj1((size_t)SG_T * npatch);
  std::vector<double> fs((size_t)SG_T * npatch), ft((size_t)SG_T * npatch);
  std::vector<std::pair<long long, int>> keys;
  int ni_max = 0, rounds_max = 0;
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
        if (pw[q % C.nrad] != 0.0) ui[i1m] = ui[(i1m + 1) % m] = 1;
      }
    }
    int ilo, ni;
    cyclic_cover(ui, m, ilo, ni);
    if (ni == m) ni = m + 1;  // a full circle needs the first row once more so that node+1 stays inside
    row_win[(size_t)tl * 2 + 0] = ilo;
    row_win[(size_t)tl * 2 + 1] = ni;
    ni_max = std::max(ni_max, ni);
    // the tile's valid patch points sorted by spline cell (phi column, theta row), ties in target-major order
    keys.clear();
    for (int w = 0; w < SG_T; w++) {
      if (row_tgt[(size_t)tl * SG_T + w] < 0) continue;
      for (int q = 0; q < npatch; q++) {
        // patch points whose quadrature weight is exactly zero (the mask table vanishes on its last interval: the
        // outermost radial node of every ray) contribute exactly zero: they are left out of the cached path
        if (pw[q % C.nrad] == 0.0) continue;
        const int wi = (ni1[(size_t)w * npatch + q] - ilo + m) % m;
        const long long cellkey = (long long)nj1[(size_t)w * npatch + q] * ni + wi;
        // per-warp mode: target first, then spline cell; tile mode: spline cell first
        const long long major = per_warp ? (long long)w * (1LL << 24) + cellkey : cellkey;
        keys.push_back({major * (long long)NPT + (w * K * 32 + q), w * npatch + q});
      }
    }
    std::sort(keys.begin(), keys.end());
    std::vector<std::vector<int2>> wchunks(SG_T);
    // chunks: runs of <= SG_PC consecutive points of one spline cell; a thread evaluates one chunk per round from
    // node data it loads once
    size_t sp = 0;
    while (sp < keys.size()) {
      const long long major = keys[sp].first / NPT;
      const long long cellkey = per_warp ? major % (1LL << 24) : major;
      int cnt = 0;
      while (sp + cnt < keys.size() && cnt < pc_max && keys[sp + cnt].first / NPT == major) cnt++;
      const int wi = (int)(cellkey % ni), j0 = (int)(cellkey / ni);
      const int2 chv = make_int2(wi | (j0 << 8) | (cnt << 18), (int)sp);
      if (per_warp)
        wchunks[keys[sp].second / npatch].push_back(chv);
      else
        chunks[tl].push_back(chv);
      for (int u = 0; u < cnt; u++) {
        const int wq = keys[sp + u].second, w = wq / npatch, q = wq - w * npatch;
        pt_dest[(size_t)tl * NPT + sp + u] = w * K * 32 + q;
        st[2 * ((size_t)tl * NPT + sp + u)] = fs[wq];
        st[2 * ((size_t)tl * NPT + sp + u) + 1] = ft[wq];
        const int k = q / 32, lane = q % 32;
        pos[(size_t)tl * NPT + (k * SG_T + w) * 32 + lane] = (int)(sp + u);
      }
      sp += cnt;
    }
    if (per_warp) {
      // warp w of the CTA owns target w: its chunks go to slots (round*SG_T + w)*32 + lane
      size_t mx = 0;
      for (int w = 0; w < SG_T; w++) mx = std::max(mx, wchunks[w].size());
      const int R = (int)((mx + 31) / 32);
      chunks[tl].assign((size_t)R * SG_T * 32, make_int2(0, 0));
      for (int w = 0; w < SG_T; w++)
        for (size_t u = 0; u < wchunks[w].size(); u++)
          chunks[tl][((u / 32) * SG_T + w) * 32 + (u % 32)] = wchunks[w][u];
    }
    row_rounds[tl] = (int)((chunks[tl].size() + sg_nt() - 1) / sg_nt());
    rounds_max = std::max(rounds_max, row_rounds[tl]);
  }
  const int CH = rounds_max * sg_nt();  // chunk slots per tile row (empty chunks: cnt = 0)
  std::vector<int2> chunk_tab((size_t)ntl * CH, make_int2(0, 0));
  for (int tl = 0; tl < ntl; tl++)
    for (size_t u = 0; u < chunks[tl].size(); u++) chunk_tab[(size_t)tl * CH + u] = chunks[tl][u];
  C.sg_ni_max = ni_max;
  C.sg_chunk_stride = CH;
  // shared memory: band of the spline (6 double2 planes x (n+1) phi columns x ni theta rows) + contribution buffer
  const size_t smem = (size_t)12 * sg_pitch_host(ni_max) * (n + 1) * sizeof(double) + (per_warp ? 0 : (size_t)3 * NPT * sizeof(double));
  if (smem > SG_SMEM_MAX || ni_max > 255) return RBC3D_OK;  // direct kernel only
  C.sg_smem = smem;
  RBC_TRY(C.sg_tile_tgt.resize(row_tgt.size()));
  RBC_TRY(C.sg_tile_win.resize(row_win.size()));
  RBC_TRY(C.sg_rounds.resize(row_rounds.size()));
  RBC_TRY(C.sg_idx.resize(pt_dest.size()));
  RBC_TRY(C.sg_pos.resize(pos.size()));
  RBC_TRY(C.sg_st.resize(st.size()));
  RBC_TRY(C.sg_chunk.resize(chunk_tab.size()));
  CUDA_TRY(cudaMemcpyAsync(C.sg_tile_tgt.p, row_tgt.data(), sizeof(int) * row_tgt.size(), cudaMemcpyHostToDevice,
                           c->stream));
  CUDA_TRY(cudaMemcpyAsync(C.sg_tile_win.p, row_win.data(), sizeof(int) * row_win.size(), cudaMemcpyHostToDevice,
                           c->stream));
  CUDA_TRY(cudaMemcpyAsync(C.sg_rounds.p, row_rounds.data(), sizeof(int) * row_rounds.size(), cudaMemcpyHostToDevice,
                           c->stream));
  CUDA_TRY(cudaMemcpyAsync(C.sg_idx.p, pt_dest.data(), sizeof(int) * pt_dest.size(), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(C.sg_pos.p, pos.data(), sizeof(int) * pos.size(), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(C.sg_st.p, st.data(), sizeof(double) * st.size(), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(C.sg_chunk.p, chunk_tab.data(), sizeof(int2) * chunk_tab.size(), cudaMemcpyHostToDevice,
                           c->stream));
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
  const int *active_list;  // [slot] -> cell
  const double *tab_dl;
  double4 *cache;  // [slot of an active cell][tile][sorted position]
};

__global__ void __launch_bounds__(SG_T * 32) k_sing_cache_build(CacheArgs a) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, slot = blockIdx.y, cell = a.active_list[slot];
  const int tl = tile / a.ntn, tn = tile - tl * a.ntn;
  const int pt0 = a.row_tgt[tl * SG_T + w];
  const int NPT = a.K * SG_T * 32;
  double4 *out = a.cache + ((size_t)slot * gridDim.x + tile) * NPT;
  const int *pos = a.pos + (size_t)tl * NPT;
  if (pt0 < 0) return;  // no target in this slot: its cache entries do not exist
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
    const int sp = pos[(k * SG_T + w) * 32 + lane];
    if (sp < 0) continue;  // beyond the patch, or a point with zero quadrature weight: not part of the cache
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
    out[sp] = r;
  }
}

// spline re-layout: ABI [cell][4 (u,u1,u2,u12)][3][n][m] -> [cell][6 planes][n][m] double2 with plane 2l = (u_l, u1_l)
// and plane 2l+1 = (u2_l, u12_l): the four Hermite data of one variable at one node are two 16-byte loads, and
// neighbouring nodes are neighbouring 16-byte words (conflict-free LDS.128 for lanes on neighbouring spline cells)
__global__ void __launch_bounds__(256) k_spline_interleave(int ncell, int plane, const double *__restrict__ sp,
                                                           double *__restrict__ out) {
  const size_t total = (size_t)ncell * plane;
  double2 *o2 = reinterpret_cast<double2 *>(out);
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t cell = e / plane, node = e - cell * plane;
    const double *src = sp + cell * 12 * (size_t)plane + node;
#pragma unroll
    for (int l = 0; l < 3; l++) {
      // a = arr*3 + var, arr = 0 (u), 1 (u1), 2 (u2), 3 (u12)
      o2[(cell * 6 + 2 * l) * plane + node] = make_double2(src[(size_t)l * plane], src[(size_t)(3 + l) * plane]);
      o2[(cell * 6 + 2 * l + 1) * plane + node] = make_double2(src[(size_t)(6 + l) * plane], src[(size_t)(9 + l) * plane]);
    }
  }
}

// pitch (in nodes) of a phi column of the shared-memory band: the smallest value >= ni with pitch = r (mod 8); the
// residue decides which neighbouring spline cells share a 16-byte bank group (RBC3D_SING_PITCH_MOD, default 1 = odd)
static int sg_pitch_mod() {
  static const int v = [] {
    const char *e = getenv("RBC3D_SING_PITCH_MOD");
    const int q = e ? atoi(e) : 1;
    return (q >= 0 && q < 8) ? q : 1;
  }();
  return v;
}
__host__ __device__ inline int sg_pitch(int ni, int r) {
  if (r == 1) return ni | 1;
  return ni + ((r - (ni & 7)) & 7);
}

static int sg_pitch_host(int ni) {  // worst case over the rows: the largest row decides the allocation
  int worst = 0;
  for (int k = 0; k <= 7 && ni - k > 0; k++) worst = worst > sg_pitch(ni - k, sg_pitch_mod()) ? worst : sg_pitch(ni - k, sg_pitch_mod());
  return worst;
}

struct BandArgs {
  int pitch_mod;
  int ncell, npc, nlat, nlon, ntl, ntn, Np, K, chunk_stride;
  const int *row_tgt;      // [tile row][T]: mesh point (ilon*nlat + ilat) of the targets of tile column 0, -1 = none
  const int *row_win;      // [tile row][2]: first theta row of the band, number of rows
  const int *row_rounds;   // [tile row]: chunk rounds (chunks / 256, rounded up)
  const int2 *chunk;       // [tile row][chunk]: (theta row | phi column << 8 | points << 18, first sorted position)
  const int *pt_dest;      // [tile row][sorted position]: target-major slot of the contribution buffer
  const double2 *pt_st;    // [tile row][sorted position]: fractional coordinates in the spline cell
  const double2 *spGp;     // [cell][6][n][m]
  const double4 *cache;    // [cell][tile][sorted position]
  const double *Bcell;
  const int *active;
  const int *active_list;  // [slot] -> cell with active targets (the cache is indexed by slot)
  double c2;
  double *acc;
};

// streaming (evict-first) 32-byte load: the cache is read exactly once per matvec
__device__ __forceinline__ double4 ld_stream4(const double4 *p) {
  double4 v;  // one 32-byte request, no L1 allocation (L1 is kept for the cell-independent tables)
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w)
               : "l"(p));
  return v;
}

// One CTA per (cell, tile row).  The band of spline(g detJ) the row's patches touch (all phi columns, ni theta rows)
// is staged once and serves the ntn tiles of the row.  A thread evaluates one chunk per round: up to SG_PC patch
// points that lie in the same spline cell, from the 48 Hermite data of the cell's four nodes loaded ONCE into
// registers (shared-memory traffic per patch point drops from 384 B to ~100 B).  The geometry cache of the next
// round is in flight while the current one is evaluated.  Contributions go to a target-major shared buffer that one
// warp per target sums in the reference's patch order.
template <bool TAB_SMEM, int SG_PC, int NT, bool PW, bool LR = false>  // tables in shared memory (when they fit) or
                                                       // through L1; NT threads; PW: warp = target, register accumulation;
                                                       // LR: low-register body (one density component at a time)
__global__ void __launch_bounds__(NT, 1) k_sing_band(BandArgs a) {
  extern __shared__ double smem[];
  // PW with more than SG_T warps: warp groups of SG_T warps take the tiles of the row in turn (no barriers in PW mode)
  constexpr int NTT = PW ? SG_T * 32 : NT, NG = NT / NTT;
  const int tid_cta = threadIdx.x, w_cta = tid_cta >> 5, lane = tid_cta & 31;
  const int tid = PW ? tid_cta % NTT : tid_cta, w = PW ? w_cta % SG_T : w_cta, grp = PW ? w_cta / SG_T : 0;
  const int slot = blockIdx.x / a.ntl, tl = blockIdx.x - slot * a.ntl;
  const int cell = a.active_list[slot];
  const int ilo = a.row_win[tl * 2 + 0], ni = a.row_win[tl * 2 + 1];
  const int m = 2 * a.nlat, n = a.nlon, plane = m * n;
  const int K = a.K, NPT = K * SG_T * 32;
  const int nip = sg_pitch(ni, a.pitch_mod);         // odd pitch of a phi column: neighbouring columns fall into
                                                     // different 16-byte bank groups
  const int wn = nip * (n + 1);                      // node slots of the band
  double2 *sP = reinterpret_cast<double2 *>(smem);   // [6][n+1][nip]
  double *sC = smem + (size_t)12 * wn;               // [3][NPT] contributions, target-major (tile mode only)
  const double hx = RBC_TWO_PI / (double)m, hy = RBC_TWO_PI / (double)n;
  const int R = a.row_rounds[tl];
  const int2 *chunk = a.chunk + (size_t)tl * a.chunk_stride + tid;
  const int *pt_dest = a.pt_dest + (size_t)tl * NPT;
  const double2 *pt_st = a.pt_st + (size_t)tl * NPT;
  if (TAB_SMEM) {  // [NPT] double2, [R*256] int2, [NPT] int behind the contribution buffer
    double2 *s_st = reinterpret_cast<double2 *>(sC + (PW ? 0 : (size_t)3 * NPT));
    int2 *s_ch = reinterpret_cast<int2 *>(s_st + NPT);
    int *s_dest = reinterpret_cast<int *>(s_ch + R * NTT);
    for (int u = tid_cta; u < NPT; u += NT) {
      s_st[u] = __ldg(pt_st + u);
      if (!PW) s_dest[u] = __ldg(pt_dest + u);
    }
    for (int u = tid_cta; u < R * NTT; u += NT) s_ch[u] = __ldg(chunk - tid + u);
    pt_st = s_st;
    pt_dest = s_dest;
    chunk = s_ch + tid;
    __syncthreads();
  }
  auto ld_tab = [&](auto *ptr) { return TAB_SMEM ? *ptr : __ldg(ptr); };
  const double4 *cg = a.cache + ((size_t)slot * a.ntl * a.ntn + (size_t)tl * a.ntn) * NPT;
  // first round of the first tile in flight before the band is staged
  int2 ch_next = ld_tab(chunk);
  double4 c_next[SG_PC];
#pragma unroll
  for (int p = 0; p < SG_PC; p++)
    c_next[p] = (p < (ch_next.x >> 18) && grp < a.ntn) ? ld_stream4(cg + (size_t)grp * NPT + ch_next.y + p)
                                                       : make_double4(0, 0, 0, 0);
  // stage the band: rows = (plane, phi column 0..n with column n = column 0), each a cyclic run of ni double2
  for (int row = w_cta; row < 6 * (n + 1); row += NT / 32) {
    const int q = row / (n + 1), wj = row - q * (n + 1);
    const int j = wj == n ? 0 : wj;
    const double2 *src = a.spGp + ((size_t)cell * 6 + q) * plane + (size_t)j * m;
    double2 *dst = sP + (size_t)q * wn + (size_t)wj * nip;
    for (int wi = lane; wi < ni; wi += 32) {
      int i = ilo + wi;
      if (i >= m) i -= m;
      dst[wi] = __ldg(src + i);
    }
  }
  if (!PW)
    for (int u = tid_cta; u < 3 * NPT; u += NT) sC[u] = 0.0;  // slots beyond npatch stay zero
  __syncthreads();
  const int pt0 = w < SG_T ? a.row_tgt[tl * SG_T + w] : -1;  // warps beyond the tile's targets only evaluate chunks
  const double c2m = a.c2 * a.Bcell[cell];  // c2Mod, ModIntOnRbcs.F90:116
  for (int tn = grp; tn < a.ntn; tn += NG) {
    const int jshift = tn * SG_TLON;
    const int ti = cell * a.npc + pt0 + tn * SG_TLON * a.nlat;
    const bool t_on = pt0 >= 0 && lane == 0 && a.active[ti] != 0;  // requested now, needed after the rounds
    double pvx = 0, pvy = 0, pvz = 0;  // PW: this lane's share of the target's sum
    for (int r = 0; r < R; r++) {
      const int2 ch = ch_next;
      double4 c4[SG_PC];
      if (LR && SG_PC > 2) {
        // no register room for a second set of cache entries: this round's entries are requested here and first used
        // after the three density components have been interpolated (16 warps cover the rest of the latency)
#pragma unroll
        for (int p = 0; p < SG_PC; p++)
          c4[p] = (p < (ch.x >> 18)) ? ld_stream4(cg + (size_t)tn * NPT + ch.y + p) : make_double4(0, 0, 0, 0);
      } else {
#pragma unroll
        for (int p = 0; p < SG_PC; p++) c4[p] = c_next[p];
      }
      {  // next round (of this tile or of the next tile of the row)
        int nr = r + 1, nt = tn;
        if (nr == R) {
          nr = 0;
          nt = tn + NG;
        }
        if (nt < a.ntn) {
          ch_next = ld_tab(chunk + nr * NTT);
          if (!(LR && SG_PC > 2)) {
            const double4 *cgn = cg + (size_t)nt * NPT + ch_next.y;
            const int cn = ch_next.x >> 18;
#pragma unroll
            for (int p = 0; p < SG_PC; p++) c_next[p] = (p < cn) ? ld_stream4(cgn + p) : make_double4(0, 0, 0, 0);
          }
        }
      }
      const int cnt = ch.x >> 18;
      if (cnt == 0) continue;
      int j = ((ch.x >> 8) & 1023) + jshift;
      if (j >= n) j -= n;
      const int a11 = j * nip + (ch.x & 255);
      // tables of the chunk's points, all requested before the first use (clamped index: no branch in the way)
      double2 stq[SG_PC];
      int destq[SG_PC];
#pragma unroll
      for (int p = 0; p < SG_PC; p++) {
        const int e = ch.y + min(p, cnt - 1);
        stq[p] = ld_tab(pt_st + e);
        destq[p] = PW ? 0 : ld_tab(pt_dest + e);
      }
      if (LR) {
        // low-register body: the Hermite data of ONE density component at a time (8 double2 instead of 24 live at
        // once), so that 512 threads fit the register file without spills and 16 warps hide the LDS / HBM latency
        double cxq[SG_PC][4], cyq[SG_PC][4], qd[SG_PC];
#pragma unroll
        for (int p = 0; p < SG_PC; p++) {
          const double s = stq[p].x, t = stq[p].y;
          cxq[p][0] = 1.0 + s * s * (-3.0 + 2.0 * s), cxq[p][1] = s * s * (3.0 - 2.0 * s);
          cxq[p][2] = hx * s * (1.0 + s * (-2.0 + s)), cxq[p][3] = hx * s * s * (-1.0 + s);
          cyq[p][0] = 1.0 + t * t * (-3.0 + 2.0 * t), cyq[p][1] = t * t * (3.0 - 2.0 * t);
          cyq[p][2] = hy * t * (1.0 + t * (-2.0 + t)), cyq[p][3] = hy * t * t * (-1.0 + t);
          qd[p] = 0.0;
        }
#pragma unroll
        for (int l = 0; l < 3; l++) {
          const double2 *P0 = sP + (size_t)(2 * l) * wn + a11, *P1 = P0 + wn;
          const double2 u11 = P0[0], u21 = P0[1], u12 = P0[nip], u22 = P0[nip + 1];
          const double2 w11 = P1[0], w21 = P1[1], w12 = P1[nip], w22 = P1[nip + 1];
#pragma unroll
          for (int p = 0; p < SG_PC; p++) {
            const double *cx = cxq[p], *cy = cyq[p];
            const double r0 = u11.x * cy[0] + u12.x * cy[1] + w11.x * cy[2] + w12.x * cy[3];
            const double r1 = u21.x * cy[0] + u22.x * cy[1] + w21.x * cy[2] + w22.x * cy[3];
            const double r2 = u11.y * cy[0] + u12.y * cy[1] + w11.y * cy[2] + w12.y * cy[3];
            const double r3 = u21.y * cy[0] + u22.y * cy[1] + w21.y * cy[2] + w22.y * cy[3];
            const double gl = cx[0] * r0 + cx[1] * r1 + cx[2] * r2 + cx[3] * r3;
            const double cl = l == 0 ? c4[p].x : l == 1 ? c4[p].y : c4[p].z;
            qd[p] = fma(cl, gl, qd[p]);
          }
          asm volatile("" ::: "memory");  // keep the next component's loads behind this component's arithmetic
        }
#pragma unroll
        for (int p = 0; p < SG_PC; p++) {
          if (p < cnt) {
            const double4 c = c4[p];
            const double q = c.w * qd[p];
            const int dest = destq[p];
            if (PW) {
              pvx += q * c.x;
              pvy += q * c.y;
              pvz += q * c.z;
            } else {
              sC[dest] = q * c.x;
              sC[NPT + dest] = q * c.y;
              sC[2 * NPT + dest] = q * c.z;
            }
          }
        }
        continue;
      }
      // Hermite data of the four nodes: nd[node][plane] = (u_l, u1_l) for plane 2l, (u2_l, u12_l) for plane 2l+1
      double2 n11[6], n21[6], n12[6], n22[6];
#pragma unroll
      for (int q = 0; q < 6; q++) {
        const double2 *P = sP + (size_t)q * wn + a11;
        n11[q] = P[0];
        n21[q] = P[1];
        n12[q] = P[nip];
        n22[q] = P[nip + 1];
      }
#pragma unroll
      for (int p = 0; p < SG_PC; p++) {
        if (p < cnt) {
          const double2 stv = stq[p];
          const int dest = destq[p];
          const double s = stv.x, t = stv.y;
          const double cx[4] = {1.0 + s * s * (-3.0 + 2.0 * s), s * s * (3.0 - 2.0 * s), hx * s * (1.0 + s * (-2.0 + s)),
                                hx * s * s * (-1.0 + s)};
          const double cy[4] = {1.0 + t * t * (-3.0 + 2.0 * t), t * t * (3.0 - 2.0 * t), hy * t * (1.0 + t * (-2.0 + t)),
                                hy * t * t * (-1.0 + t)};
          double g[3];
#pragma unroll
          for (int l = 0; l < 3; l++) {
            // same association as spline_interp (device_math.cuh): U = .x of plane 2l, U1 = .y, U2 = .x of 2l+1, U12 = .y
            const double r0 = n11[2 * l].x * cy[0] + n12[2 * l].x * cy[1] + n11[2 * l + 1].x * cy[2] + n12[2 * l + 1].x * cy[3];
            const double r1 = n21[2 * l].x * cy[0] + n22[2 * l].x * cy[1] + n21[2 * l + 1].x * cy[2] + n22[2 * l + 1].x * cy[3];
            const double r2 = n11[2 * l].y * cy[0] + n12[2 * l].y * cy[1] + n11[2 * l + 1].y * cy[2] + n12[2 * l + 1].y * cy[3];
            const double r3 = n21[2 * l].y * cy[0] + n22[2 * l].y * cy[1] + n21[2 * l + 1].y * cy[2] + n22[2 * l + 1].y * cy[3];
            g[l] = cx[0] * r0 + cx[1] * r1 + cx[2] * r2 + cx[3] * r3;
          }
          const double4 c = c4[p];
          const double qd = c.w * (c.x * g[0] + c.y * g[1] + c.z * g[2]);
          if (PW) {
            pvx += qd * c.x;
            pvy += qd * c.y;
            pvz += qd * c.z;
          } else {
            sC[dest] = qd * c.x;
            sC[NPT + dest] = qd * c.y;
            sC[2 * NPT + dest] = qd * c.z;
          }
        }
      }
    }
    if (PW) {  // fixed order: chunk order per lane, then the shuffle tree -- no barrier, warps run independently
      pvx = warp_sum(pvx);
      pvy = warp_sum(pvy);
      pvz = warp_sum(pvz);
      if (t_on) {
        atomicAdd(a.acc + ti, c2m * pvx);
        atomicAdd(a.acc + (size_t)a.Np + ti, c2m * pvy);
        atomicAdd(a.acc + 2 * (size_t)a.Np + ti, c2m * pvz);
      }
      continue;
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
      if (t_on) {  // single writer per address: a reduction without return value, no load to wait for
        atomicAdd(a.acc + ti, c2m * dvx);
        atomicAdd(a.acc + (size_t)a.Np + ti, c2m * dvy);
        atomicAdd(a.acc + 2 * (size_t)a.Np + ti, c2m * dvz);
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
  // compact list of the cells this rank owns targets of: the singular cache and its kernel only cover those
  std::vector<int> flags(C.ncell), list;
  CUDA_TRY(cudaMemcpyAsync(flags.data(), C.sg_cell_active.p, sizeof(int) * C.ncell, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < C.ncell; i++)
    if (flags[i]) list.push_back(i);
  C.sg_nactive = (int)list.size();
  RBC_TRY(C.sg_active_list.resize(list.size() > 0 ? list.size() : 1));
  if (!list.empty())
    CUDA_TRY(cudaMemcpyAsync(C.sg_active_list.p, list.data(), sizeof(int) * list.size(), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RBC3D_OK;
}

// geometry time (SourceList_UpdateCoord): density-independent cache of the double-layer patch integrand
int singular_prepare(rbc3d_ctx *c) {
  Cells &C = c->cells;
  C.sg_cache_ok = false;
  if (!C.sg_ok || C.Np == 0 || c->sing_cache_mode == 0) return RBC3D_OK;
  const size_t per_cell = (size_t)C.sg_ntiles * C.sg_K * SG_T * 32;
  const size_t ncache = (size_t)std::max(C.sg_nactive, 1);
  const size_t need = per_cell * ncache * sizeof(double4);
  if (C.sg_cache.n < per_cell * ncache) {
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    // leave room for the interleaved density spline and the rest of the working set
    const size_t reserve = (size_t)C.ncell * 12 * 2 * C.nlat * C.nlon * 8 * 2 + ((size_t)2 << 30);
    if (need + reserve > free_b + C.sg_cache.n * sizeof(double4)) return RBC3D_OK;  // direct kernel
    if (C.sg_cache.resize(per_cell * ncache) != RBC3D_OK) return RBC3D_OK;
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
  a.active_list = C.sg_active_list.p;
  // 65535 limit of gridDim.y: chunk the active cells
  for (int c0 = 0; c0 < C.sg_nactive; c0 += 32768) {
    CacheArgs b = a;
    const int nc = std::min(32768, C.sg_nactive - c0);
    b.active_list = a.active_list + c0;
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
  a.pitch_mod = sg_pitch_mod();
  a.ncell = C.ncell;
  a.npc = C.npc;
  a.nlat = C.nlat;
  a.nlon = C.nlon;
  a.ntl = C.sg_ntl;
  a.ntn = C.sg_ntn;
  a.Np = C.Np;
  a.K = C.sg_K;
  a.chunk_stride = C.sg_chunk_stride;
  a.row_tgt = C.sg_tile_tgt.p;
  a.row_win = C.sg_tile_win.p;
  a.row_rounds = C.sg_rounds.p;
  a.chunk = C.sg_chunk.p;
  a.pt_dest = C.sg_idx.p;
  a.pt_st = reinterpret_cast<const double2 *>(C.sg_st.p);
  a.spGp = reinterpret_cast<const double2 *>(C.spGi.p);
  a.cache = C.sg_cache.p;
  a.Bcell = C.B.p;
  a.active = t.active.p;
  a.active_list = C.sg_active_list.p;
  a.c2 = c2;
  a.acc = t.acc.p;
  const int grid = C.sg_nactive * C.sg_ntl;
  if (grid == 0) return RBC3D_OK;
  // tables of a tile row in shared memory when they fit behind the band and the contribution buffer
  const size_t NPT = (size_t)C.sg_K * SG_T * 32;
  const size_t tab = NPT * (sizeof(double2) + sizeof(int)) + (size_t)C.sg_chunk_stride * sizeof(int2);
  static const bool no_tab = getenv("RBC3D_SING_TAB_GLOBAL") != nullptr;
  const bool tab_smem = C.sg_smem + tab <= SG_SMEM_MAX && !no_tab;
  const size_t smem = C.sg_smem + (tab_smem ? tab : 0);
#define LAUNCH_BAND(TS_, PC_, NT_, PW_)                                                                                       \
  do {                                                                                                                        \
    CUDA_TRY(cudaFuncSetAttribute(k_sing_band<TS_, PC_, NT_, PW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_sing_band<TS_, PC_, NT_, PW_><<<grid, NT_, smem, c->stream>>>(a);                                                      \
  } while (0)
#define LAUNCH_BAND_NT(TS_, PC_)                                \
  do {                                                          \
    if (sg_lr() == 1 && PC_ == 2) {                                                                                            \
      CUDA_TRY(cudaFuncSetAttribute(k_sing_band<TS_, 2, 512, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      k_sing_band<TS_, 2, 512, false, true><<<grid, 512, smem, c->stream>>>(a);                                               \
    } else if (sg_lr() == 2 && PC_ <= 4) {                                                                                     \
      CUDA_TRY(cudaFuncSetAttribute(k_sing_band<TS_, PC_, 512, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      k_sing_band<TS_, PC_, 512, true, true><<<grid, 512, smem, c->stream>>>(a);                                              \
    } else                                                      \
    if (sg_per_warp()) LAUNCH_BAND(TS_, PC_, SG_T * 32, true);  \
    else if (sg_nt() == 512) LAUNCH_BAND(TS_, PC_, 512, false); \
    else if (sg_nt() == 384) LAUNCH_BAND(TS_, PC_, 384, false); \
    else LAUNCH_BAND(TS_, PC_, 256, false);                     \
  } while (0)
  switch (sg_pc() * 2 + (tab_smem ? 1 : 0)) {
    case 4: LAUNCH_BAND_NT(false, 2); break;
    case 5: LAUNCH_BAND_NT(true, 2); break;
    case 6: LAUNCH_BAND_NT(false, 3); break;
    case 7: LAUNCH_BAND_NT(true, 3); break;
    case 8: LAUNCH_BAND_NT(false, 4); break;
    case 9: LAUNCH_BAND_NT(true, 4); break;
    case 12: LAUNCH_BAND_NT(false, 6); break;
    default: LAUNCH_BAND_NT(true, 6); break;
  }
#undef LAUNCH_BAND_NT
#undef LAUNCH_BAND
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
