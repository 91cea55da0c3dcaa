// singular.cu -- singular self-cell correction of the real-space sum: RBC_SingInt (ModRbcSingInt.F90:29-90),
// called from AddIntOnRbcs for every on-surface target (ModIntOnRbcs.F90:114-121).
//
// For a target (cell, ilat0, ilon0) the masked-out part of the pair sum is replaced by quadrature on a polar
// patch of nrad x nazm points whose reference-sphere coordinates thG/phiG are shared by all cells; positions,
// normals and densities at the patch points come from the cell's bicubic splines (Spline_Interp,
// ModSpline.F90:150-191).
//
// Two paths:
//  * k_singular (direct): one warp per target, everything evaluated from the ABI-layout splines.  Used whenever the
//    geometry cache is not available (no memory, a mesh the row kernel does not cover, sing_cache_mode 0).
//  * k_sing_row (cached geometry): the patch of (ilat, ilon) is the patch of (ilat, 0) rotated by phi(ilon) about the
//    polar axis (PolarPatch_Build, ModPolarPatch.F90:99-148) and the spline's phi nodes ARE the mesh longitudes, so
//    patch point p of target ilon lies in spline cell (i1_p, j1_p + ilon) with the SAME fractional coordinates for every
//    target of a latitude row.  Hence:
//      - LANE = TARGET (31 consecutive longitudes per warp + one lane for the right-hand column of the last target);
//        all lanes walk the row's patch points together.  Bicubic basis values, spline-cell indices and quadrature
//        weights are warp-uniform table entries (shared memory broadcasts), node data are conflict-free LDS.128 of
//        consecutive phi columns, and no cross-lane reduction is needed: a lane accumulates its own target.
//      - a lane loads only ITS phi column (2 theta nodes x 3 variables x 4 Hermite data = 24 doubles), forms the theta
//        interpolants P, Q of that column and hands P cy1 + Q cy3 (3 doubles) to its left neighbour by shuffle: half the
//        shared-memory traffic and 36 instead of 60 FMAs per 3-variable interpolation.
//      - patch points are sorted by spline cell, columns first: consecutive cells of a column keep the lower theta nodes
//        in registers ("slide").
//      - everything that does not depend on the density is cached per geometry: xx = x(patch point) - x(target) and
//        w = weight * EwaldCoeff_DL(|xx|) * (xx . a3), 32 B per patch point in two 16-byte halves [item][point][half]
//        [target], streamed once per matvec with coalesced 16-byte loads, 3 records in flight per lane.
//      - one persistent CTA per (latitude row, replica) keeps the row's tables in shared memory and walks over the
//        cells; the band of the node-interleaved spline a row's patches touch ([6 planes][theta window][nlon] double2,
//        97 KB) is double-buffered: the next cell's band arrives by cp.async.bulk (TMA) while the current one is
//        evaluated.  CTAs of all rows sweep the cells in step, so a cell's spline is read from HBM once.
//      - per-target sums are combined over the 4 point streams in a fixed order and stored by their single owner:
//        deterministic, no atomics.
//    The same kernel builds the cache (two passes: positions, then normals) and evaluates the single-layer singular
//    integrals of Compute_Rhs from the cached xx (EwaldCoeff_SL from the table on the fly).
#include <algorithm>
#include <cstdlib>

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


// ---------------------------------------------------------------------------------------------------------
// row kernel: tables (host, once per rbc3d_cells_set_mesh)

constexpr int SR_TPW = 31;        // targets per warp; lane nt (<= 31) holds the right-hand column of the last target
constexpr int SR_PF = 2;          // cache records in flight per lane (registers)
constexpr int SR_U = 2;           // points per trip of the unrolled loop; streams are padded to multiples of it
constexpr int SR_L2PF = 20;       // patch points ahead of which the records are prefetched into L2
constexpr int SR_TABW = 10;       // doubles per table entry: cx[4], cy[4], quadrature weight, code
constexpr int SR_NT = 480;        // consumer threads per CTA at most (+ one producer warp = 16 warps of 128 registers)
constexpr int SR_FRESH = 1 << 16, SR_DUMMY = 1 << 17, SR_LOADX = 1 << 18, SR_LOADY = 1 << 19;
constexpr int SR_RI = 16;         // ints per row-info record: ilo, ni, table entries, stream bounds [0..NS] (table indices)
constexpr size_t SR_SMEM_MAX = 227 * 1024;
enum { SR_BUILD_X = 0, SR_BUILD_N = 1, SR_DL = 2, SR_SL = 3 };

static int sr_groups(int nlon) { return (nlon + SR_TPW - 1) / SR_TPW; }
static int sr_tpw(int nlon) { return (nlon + sr_groups(nlon) - 1) / sr_groups(nlon); }  // balanced groups: 72 -> 3 x 24
static int sr_streams(int nlon) {
  const int g = sr_groups(nlon);
  return std::max(1, std::min(5, (SR_NT / 32) / g));
}

int singular_mesh_prepare(rbc3d_ctx *c, const double *thG, const double *phiG, const double *pw) {
  Cells &C = c->cells;
  C.sg_ok = false;
  C.sg_cache_ok = false;
  C.spGi_valid = false;
  C.spFi_valid = false;
  const int nlat = C.nlat, nlon = C.nlon, m = 2 * nlat, n = nlon, npatch = C.nrad * C.nazm;
  const double hx = RBC_TWO_PI / (double)m, hy = RBC_TWO_PI / (double)n;
  const double ihx = 1.0 / hx, ihy = 1.0 / hy;
  C.sg_npatch_active = 0;
  for (int q = 0; q < npatch; q++) C.sg_npatch_active += (pw[q % C.nrad] != 0.0) ? 1 : 0;
  const int npts = C.sg_npatch_active;
  const int ngrp = sr_groups(n), NS = sr_streams(n);
  if (npts == 0 || n > 255 || m > 255 || ngrp * NS * 32 > SR_NT || NS + 4 > SR_RI) return RBC3D_OK;  // direct kernel only
  std::vector<std::vector<double>> rowtab(nlat);
  std::vector<int> rowinfo((size_t)nlat * SR_RI, 0);
  int ni_max = 0, ntab = 0;
  struct Pt {
    int j1, wi, q;
    double s, t;
  };
  std::vector<Pt> pts;
  std::vector<int> i1v(npatch), j1v(npatch);
  for (int row = 0; row < nlat; row++) {
    // the patch of (ilat = row, ilon = 0); thG does not depend on phi0 and phiG = atan2(..) + phi0
    const size_t p0 = (size_t)row * npatch;  // point (ilon 0, ilat row) = 0 * nlat + row
    std::vector<char> ui(m, 0);
    std::vector<double> fs(npatch), ft(npatch);
    for (int q = 0; q < npatch; q++) {
      // same arithmetic as spline_interp (device_math.cuh)
      const double xs = thG[p0 + q] * ihx, ys = phiG[p0 + q] * ihy;
      const int i1 = (int)floor(xs), j1 = (int)floor(ys);
      fs[q] = xs - (double)i1;
      ft[q] = ys - (double)j1;
      i1v[q] = ((i1 % m) + m) % m;
      j1v[q] = ((j1 % n) + n) % n;
      // patch points whose quadrature weight is exactly zero (the mask table vanishes on its last interval: the
      // outermost radial node of every ray) contribute exactly zero: they are left out of tables and cache
      if (pw[q % C.nrad] != 0.0) ui[i1v[q]] = ui[(i1v[q] + 1) % m] = 1;
    }
    int ilo, ni;
    cyclic_cover(ui, m, ilo, ni);
    if (ni >= m) return RBC3D_OK;  // a full circle of theta nodes: not a sphere patch; direct kernel only
    ni_max = std::max(ni_max, ni);
    pts.clear();
    for (int q = 0; q < npatch; q++)
      if (pw[q % C.nrad] != 0.0) pts.push_back({j1v[q], (i1v[q] - ilo + m) % m, q, fs[q], ft[q]});
    std::sort(pts.begin(), pts.end(), [](const Pt &a, const Pt &b) {
      if (a.j1 != b.j1) return a.j1 < b.j1;
      if (a.wi != b.wi) return a.wi < b.wi;
      return a.q < b.q;
    });
    // code FRESH: first point of a spline cell = load both theta nodes of the lane's column
    std::vector<int> code(npts, 0);
    std::vector<double> cost(npts + 1, 0.0);
    for (int k = 0; k < npts; k++) {
      const int f = (k == 0 || pts[k].j1 != pts[k - 1].j1 || pts[k].wi != pts[k - 1].wi) ? SR_FRESH : 0;
      code[k] = pts[k].wi | (pts[k].j1 << 8) | f;
      cost[k + 1] = cost[k] + 90.0 + (f ? 14.0 : 0.0);  // instructions per point; node loads at the first point of a spline cell
    }
    // NS streams of contiguous points of equal length (+-1): every warp of a target group then runs the same number of
    // steps; a cut inside a spline cell only costs the next stream one more node load
    std::vector<int> cut(NS + 1, 0);
    for (int s = 1; s <= NS; s++) cut[s] = (int)(((long long)npts * s) / NS);
    (void)cost;
    // table entries: every stream padded to a multiple of SR_U points with dummies (zero basis: they add exactly zero;
    // their record index is the stream's last point), SR_PF more dummies behind the row for the look-ahead
    int *ri = rowinfo.data() + (size_t)row * SR_RI;
    ri[0] = ilo, ri[1] = ni;
    std::vector<double> &rt = rowtab[row];
    rt.clear();
    // Node registers X, Y of a lane hold two adjacent theta rows of its phi column.  FRESH loads both (X = upper row i1,
    // Y = lower row i2).  When the next spline cell is the one below in the same column, only the new lower row is
    // loaded, into the register pair that held the old upper row (LOADX / LOADY), and the roles of X and Y are swapped
    // by swapping the theta basis values IN THE TABLE (te[0] <-> te[1], te[2] <-> te[3]): no register moves.
    bool swapped = false;
    auto push = [&](int k, bool dummy, int load) {
      double te[SR_TABW];
      for (int i = 0; i < SR_TABW; i++) te[i] = 0.0;
      int cw = (code[k] & 0xffff) | load;
      if (!dummy) {
        const double s = pts[k].s, t = pts[k].t;
        te[0] = 1.0 + s * s * (-3.0 + 2.0 * s), te[1] = s * s * (3.0 - 2.0 * s);
        te[2] = hx * s * (1.0 + s * (-2.0 + s)), te[3] = hx * s * s * (-1.0 + s);
        te[4] = 1.0 + t * t * (-3.0 + 2.0 * t), te[5] = t * t * (3.0 - 2.0 * t);
        te[6] = hy * t * (1.0 + t * (-2.0 + t)), te[7] = hy * t * t * (-1.0 + t);
        te[8] = pw[pts[k].q % C.nrad];
        if (swapped) std::swap(te[0], te[1]), std::swap(te[2], te[3]);
      } else {
        cw |= SR_DUMMY;
      }
      const unsigned long long w64 = (unsigned long long)(unsigned)cw | ((unsigned long long)(unsigned)k << 32);
      memcpy(&te[9], &w64, sizeof(double));
      rt.insert(rt.end(), te, te + SR_TABW);
    };
    for (int s = 0; s < NS; s++) {
      ri[3 + s] = (int)(rt.size() / SR_TABW);
      for (int k = cut[s]; k < cut[s + 1]; k++) {
        int load = 0;
        if (k == cut[s]) {
          load = SR_FRESH, swapped = false;  // a stream starts with a full node load
        } else if (code[k] & SR_FRESH) {
          if (pts[k].j1 == pts[k - 1].j1 && pts[k].wi == pts[k - 1].wi + 1) {
            load = swapped ? SR_LOADY : SR_LOADX;  // the pair that held the old upper row takes the new lower row
            swapped = !swapped;
          } else {
            load = SR_FRESH, swapped = false;
          }
        }
        push(k, false, load);
      }
      const int len = cut[s + 1] - cut[s];
      for (int k = len; k % SR_U != 0; k++) push(std::max(cut[s + 1] - 1, 0), true, 0);
    }
    ri[3 + NS] = (int)(rt.size() / SR_TABW);
    for (int k = 0; k < SR_PF; k++) push(npts - 1, true, 0);
    ri[2] = (int)(rt.size() / SR_TABW);
    ntab = std::max(ntab, ri[2]);
  }
  std::vector<double> tab((size_t)nlat * ntab * SR_TABW, 0.0);
  for (int row = 0; row < nlat; row++) std::copy(rowtab[row].begin(), rowtab[row].end(), tab.begin() + (size_t)row * ntab * SR_TABW);
  C.sg_ntab = ntab;
  C.sg_ni_max = ni_max;
  RBC_TRY(C.sg_st.resize(tab.size()));
  RBC_TRY(C.sg_idx.resize(rowinfo.size()));
  CUDA_TRY(cudaMemcpyAsync(C.sg_st.p, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(C.sg_idx.p, rowinfo.data(), sizeof(int) * rowinfo.size(), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  // shared memory: barriers, the row's tables, the cross-stream reduction buffer, one or two bands
  const size_t band = (size_t)6 * ni_max * n * sizeof(double2);
  const size_t fixed = 64 + (size_t)ntab * SR_TABW * sizeof(double) +
                       (((size_t)2 * ngrp * (NS - 1) * 3 * sr_tpw(n) + 1) & ~(size_t)1) * sizeof(double);
  if (fixed + band > SR_SMEM_MAX) return RBC3D_OK;  // direct kernel only
  C.sg_K = (fixed + 2 * band <= SR_SMEM_MAX) ? 2 : 1;  // band buffers
  C.sg_smem = fixed + C.sg_K * band;
  C.sg_ok = true;
  return RBC3D_OK;
}

// ---------------------------------------------------------------------------------------------------------
// node-interleaved spline planes: ABI [cell][4 (u,u1,u2,u12)][3][n][m] -> [cell][6][m][n] double2, plane 2l = (u_l, u1_l),
// plane 2l+1 = (u2_l, u12_l), PHI FASTEST: the lanes of a warp (consecutive target longitudes) read consecutive 16-byte
// words, and the theta window of a latitude row is one contiguous block per plane (one bulk copy)
__global__ void __launch_bounds__(256) k_spline_planes(int ncell, int m, int n, const int *__restrict__ need,
                                                       const double *__restrict__ sp, double2 *__restrict__ out) {
  const size_t plane = (size_t)m * n, total = (size_t)ncell * plane;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t cell = e / plane, node = e - cell * plane;
    if (need && !need[cell]) continue;
    const int i = (int)(node / n), j = (int)(node - (size_t)i * n);
    const double *src = sp + cell * 12 * plane + (size_t)j * m + i;
#pragma unroll
    for (int l = 0; l < 3; l++) {
      out[(cell * 6 + 2 * l) * plane + node] = make_double2(src[(size_t)l * plane], src[(size_t)(3 + l) * plane]);
      out[(cell * 6 + 2 * l + 1) * plane + node] = make_double2(src[(size_t)(6 + l) * plane], src[(size_t)(9 + l) * plane]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
struct RowArgs {
  Params prm;
  int npc, nlat, nlon, Np, npts, ntab, ni_max, nslot, ngrp, NS, tpw, reps, nbuf;
  const double *tab;        // [row][ntab][SR_TABW]
  const int *rowinfo;       // [row][SR_RI]
  const double2 *planes;    // [cell][6][m][n] of the interpolated field (x, a3, g detJ or f detJ)
  double2 *cache;           // [slot][row][point][half][n]: (xx.x, xx.y), (xx.z, w)
  const double *spx;        // ABI-layout spline of x (BUILD_X: target positions, ModRbcSingInt.F90:58)
  const double *th, *phi;
  const double *Bcell;
  const int *active;
  const int *active_list;   // [slot] -> cell
  const double *tab_sl, *tab_dl;
  double coef;              // c2 (DL) / c1 (SL)
  double *acc;              // SoA(3,Np)
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// contiguous bytes (a multiple of 16) on their way from HBM to L2, no register, no shared memory
__device__ __forceinline__ void l2_prefetch(const void *p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned long long *bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)  // suspend-time hint in ns: the producer lane sleeps in hardware
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// streaming (evict-first) 16-byte load: the cache is read exactly once per matvec; no L1 allocation
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ double2 ld_stream2(const double2 *p, unsigned long long pol) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;"
               : "=d"(v.x), "=d"(v.y)
               : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ double2 ld_plain2(const double2 *p) {  // same, for records this kernel rewrites later
  double2 v;
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_stream2(double2 *p, double2 v) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// One persistent CTA per (latitude row, replica): consumer warps = target groups x point streams, plus one producer warp
// that keeps the spline bands coming (full / empty mbarriers per band buffer: a consumer warp never waits for another
// one except for its own group's 4 streams when their sums are combined).
template <int MODE>
__global__ void __launch_bounds__(SR_NT + 32, 1) k_sing_row(RowArgs a) {  // 13 warps: 4 on one scheduler -> 128 registers
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = a.nlon, m = 2 * a.nlat, npts = a.npts, NS = a.NS;
  const int row = blockIdx.x % a.nlat, rep = blockIdx.x / a.nlat;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncons = a.ngrp * NS;                                        // consumer warps; warp ncons = producer
  unsigned long long *bar_full = reinterpret_cast<unsigned long long *>(smem_raw);   // [2]
  unsigned long long *bar_empty = bar_full + 2;                                          // [2]
  double *s_tab = reinterpret_cast<double *>(smem_raw + 64);
  double *s_red = s_tab + (size_t)a.ntab * SR_TABW;                     // [2][consumer warp][3][tpw]
  const int tpw = a.tpw;
  double2 *s_band = reinterpret_cast<double2 *>(s_red + (((size_t)2 * a.ngrp * (NS - 1) * 3 * tpw + 1) & ~(size_t)1));
  const int *ri = a.rowinfo + (size_t)row * SR_RI;
  const int ilo = ri[0], ni = ri[1];
  const int wpl = a.ni_max * n;                                         // double2 per plane of a band buffer (uniform)
  const size_t band_n = (size_t)6 * wpl;
  const int nbuf = a.nbuf;
  // ---- prologue: tables of the row, barriers ----
  for (int e = threadIdx.x; e < a.ntab * SR_TABW; e += blockDim.x) s_tab[e] = a.tab[(size_t)row * a.ntab * SR_TABW + e];
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; b++) {
      mbar_init(bar_full + b, 1);
      mbar_init(bar_empty + b, ncons);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == ncons) {
    // ---- producer: one lane streams the band of every item into the buffer ring ----
    if (lane == 0) {
      const unsigned band_bytes = (unsigned)(6 * ni * n * sizeof(double2));
      const int n1 = min(ni, m - ilo);  // rows before the cyclic wrap
      int it = 0;
      for (int slot = rep; slot < a.nslot; slot += a.reps, it++) {
        const int buf = it % nbuf, use = it / nbuf;
        if (use > 0)  // every consumer warp has left the buffer; poll with a pause: the issue slots belong to the consumers
          while (!mbar_test(bar_empty + buf, (unsigned)((use - 1) & 1))) {
          }
        const int cell = a.active_list[slot];
        mbar_expect_tx(bar_full + buf, band_bytes);
        for (int q = 0; q < 6; q++) {
          const double2 *src = a.planes + ((size_t)cell * 6 + q) * m * n;
          double2 *dst = s_band + (size_t)buf * band_n + (size_t)q * wpl;
          bulk_g2s(dst, src + (size_t)ilo * n, (unsigned)(n1 * n * sizeof(double2)), bar_full + buf);
          if (n1 < ni) bulk_g2s(dst + (size_t)n1 * n, src, (unsigned)((ni - n1) * n * sizeof(double2)), bar_full + buf);
        }
      }
    }
    return;
  }
  const int grp = warp / NS, strm = warp - grp * NS;
  const int jt0 = grp * tpw, nt = min(tpw, n - jt0);  // this warp's targets jt0 .. jt0 + nt - 1
  const bool is_t = lane < nt;
  const int lane_t = min(lane, nt - 1);               // lanes beyond the targets repeat the last target's record loads
  const int cb = (jt0 + lane) % n;                    // phi column of this lane for j1 = 0
  const int pbeg = ri[3 + strm], pend = ri[4 + strm]; // table entries of this stream, a multiple of SR_U
  const unsigned long long pol = l2_evict_first_policy();
  const size_t rstep = (size_t)2 * n;                 // double2 per patch point of an item
  const size_t item_step = (size_t)a.reps * a.nlat * npts * rstep;
  // the stream's real points are contiguous in the cache: [rfirst, rfirst + rcount)
  const int rfirst = (int)(reinterpret_cast<const unsigned long long *>(s_tab + (size_t)pbeg * SR_TABW + 9)[0] >> 32);
  const int rcount = (pend > pbeg)
                         ? (int)(reinterpret_cast<const unsigned long long *>(s_tab + (size_t)(pend - 1) * SR_TABW + 9)[0] >> 32) + 1 - rfirst
                         : 0;
  const bool pf_lane = (MODE != SR_BUILD_X) && grp == 0 && lane == 0;  // group 0 prefetches whole rows (all groups' records)
  int it = 0;
  for (int slot = rep; slot < a.nslot; slot += a.reps, it++) {
    const int buf = it % nbuf, use = it / nbuf;
    const int cell = a.active_list[slot];
    double2 *item = a.cache + (((size_t)slot * a.nlat + row) * npts) * rstep;
    double2 *rec = item + (jt0 + lane_t);
    double xi0 = 0, xi1 = 0, xi2 = 0;
    if (MODE == SR_BUILD_X && is_t) {
      double xi[3];
      spline_interp<3>(a.spx + (size_t)12 * m * n * cell, m, n, a.th[row], a.phi[jt0 + lane], xi);  // ModRbcSingInt.F90:58
      xi0 = xi[0], xi1 = xi[1], xi2 = xi[2];
    }
    if (pf_lane) {
      // HBM -> L2, far ahead of the register loads: this item's first records (first item only) and the next item's
      const unsigned head = (unsigned)(min(rcount, SR_L2PF) * rstep * sizeof(double2));
      if (head) {
        if (it == 0) l2_prefetch(item + (size_t)rfirst * rstep, head);
        if (slot + a.reps < a.nslot) l2_prefetch(item + item_step + (size_t)rfirst * rstep, head);
      }
    }
    // records: SR_PF points in flight in registers
    unsigned long long cw[SR_PF];
    double2 ra[SR_PF], rb[SR_PF];
#pragma unroll
    for (int k = 0; k < SR_PF; k++) {
      cw[k] = reinterpret_cast<const unsigned long long *>(s_tab + (size_t)(pbeg + k) * SR_TABW + 9)[0];
      ra[k] = rb[k] = make_double2(0, 0);
      if (MODE != SR_BUILD_X) {
        const double2 *q = rec + (size_t)(cw[k] >> 32) * rstep;
        ra[k] = MODE == SR_BUILD_N ? ld_plain2(q) : ld_stream2(q, pol);
        rb[k] = MODE == SR_BUILD_N ? ld_plain2(q + n) : ld_stream2(q + n, pol);
      }
    }
    mbar_wait(bar_full + buf, (unsigned)(use & 1));
    const double2 *band = s_band + (size_t)buf * band_n;
    double2 top[6], bot[6];
#pragma unroll
    for (int q = 0; q < 6; q++) top[q] = bot[q] = make_double2(0, 0);
    double pv0 = 0, pv1 = 0, pv2 = 0;
    const double *te = s_tab + (size_t)pbeg * SR_TABW;
    for (int p = pbeg; p < pend; p += SR_U) {
      if (pf_lane) {
        const int r0 = (int)(cw[0] >> 32) + SR_L2PF;  // records SR_L2PF points ahead, SR_U points per trip
        if (r0 + SR_U <= rfirst + rcount) l2_prefetch(item + (size_t)r0 * rstep, (unsigned)(SR_U * rstep * sizeof(double2)));
      }
#pragma unroll
      for (int k = 0; k < SR_U; k++) {
        const int kr = k % SR_PF;
        const int code = (int)(unsigned)cw[kr];
        if (code & (SR_FRESH | SR_LOADX | SR_LOADY)) {  // first point of a spline cell: theta nodes of this lane's phi column
          int col = ((code >> 8) & 255) + cb;
          if (col >= n) col -= n;
          const double2 *nb = band + (code & 255) * n + col;
          if (code & SR_FRESH) {
#pragma unroll
            for (int q = 0; q < 6; q++) top[q] = nb[q * wpl];
          }
          if (code & (SR_FRESH | SR_LOADY)) {
#pragma unroll
            for (int q = 0; q < 6; q++) bot[q] = nb[q * wpl + n];
          }
          if (code & SR_LOADX) {  // the cell below: its lower row replaces the old upper row (roles swapped in the table)
#pragma unroll
            for (int q = 0; q < 6; q++) top[q] = nb[q * wpl + n];
          }
        }
        const double2 cx01 = *reinterpret_cast<const double2 *>(te), cx23 = *reinterpret_cast<const double2 *>(te + 2);
        const double2 cy01 = *reinterpret_cast<const double2 *>(te + 4), cy23 = *reinterpret_cast<const double2 *>(te + 6);
        double g[3];
#pragma unroll
        for (int l = 0; l < 3; l++) {
          // theta interpolants of this lane's phi column: P (value), Q (phi derivative)
          const double P = top[2 * l].x * cx01.x + bot[2 * l].x * cx01.y + top[2 * l].y * cx23.x + bot[2 * l].y * cx23.y;
          const double Q = top[2 * l + 1].x * cx01.x + bot[2 * l + 1].x * cx01.y + top[2 * l + 1].y * cx23.x +
                           bot[2 * l + 1].y * cx23.y;
          const double A = P * cy01.x + Q * cy23.x;   // this column as the left one of the lane's own target
          const double Br = P * cy01.y + Q * cy23.y;  // ... as the right one of the left neighbour's target
          g[l] = A + __shfl_down_sync(FULL_MASK, Br, 1);
        }
        const double2 A2 = ra[kr], B2 = rb[kr];
        double2 *rp = rec + (size_t)(cw[kr] >> 32) * rstep;  // this point's records (stores of the build passes)
        if (MODE == SR_BUILD_X) {
          if (is_t && !(code & SR_DUMMY)) {
            st_stream2(rp, make_double2(g[0] - xi0, g[1] - xi1));
            st_stream2(rp + n, make_double2(g[2] - xi2, 0.0));
          }
        } else if (MODE == SR_BUILD_N) {
          const double xx = A2.x, yy = A2.y, zz = B2.x;
          const double r2 = xx * xx + yy * yy + zz * zz;
          double w = 0.0;
          // rr < rc (ModRbcSingInt.F90:69) and EwaldCoeff_DL (ModEwaldFunc.F90:141-178: zero below r_eps and beyond the
          // table) without divisions: 1/r from the hardware seed + one cubic step (rsqrt_pos), as the pair cache does
          if (r2 >= a.prm.r_eps * a.prm.r_eps && !(r2 > a.prm.rc2_thr)) {
            const double rinv = rsqrt_pos(r2), rr = r2 * rinv;
            if (rr < a.prm.rc) {
              const double sc = rr * a.prm.tab_scale;
              const int i0 = (int)sc;
              if (i0 < RBC3D_NTAB) {
                const double t0 = __ldg(a.tab_dl + i0), t1 = __ldg(a.tab_dl + i0 + 1);
                const double cdl = t0 * ((double)(i0 + 1) - sc) + t1 * (sc - (double)i0);
                const double ir2 = rinv * rinv;
                w = cdl * ir2 * ir2 * rinv * te[8] * (xx * g[0] + yy * g[1] + zz * g[2]);
              }
            }
          }
          if (is_t && !(code & SR_DUMMY)) st_stream2(rp + n, make_double2(zz, w));
        } else if (MODE == SR_DL) {
          const double qd = B2.y * (A2.x * g[0] + A2.y * g[1] + B2.x * g[2]);
          pv0 = fma(qd, A2.x, pv0);
          pv1 = fma(qd, A2.y, pv1);
          pv2 = fma(qd, B2.x, pv2);
        } else {  // SR_SL: ModRbcSingInt.F90:72-78 with the cached xx
          const double xx = A2.x, yy = A2.y, zz = B2.x;
          const double rr = sqrt(xx * xx + yy * yy + zz * zz);
          if (rr < a.prm.rc) {
            const double wq = te[8];
            const double f0 = g[0] * wq, f1 = g[1] * wq, f2 = g[2] * wq;
            double EA, EB;
            ewald_sl(a.tab_sl, a.prm, rr, EA, EB);
            const double xf = EA * (xx * f0 + yy * f1 + zz * f2);
            pv0 += xf * xx + EB * f0;
            pv1 += xf * yy + EB * f1;
            pv2 += xf * zz + EB * f2;
          }
        }
        // look ahead: the table has SR_PF dummy entries behind the last stream
        cw[kr] = reinterpret_cast<const unsigned long long *>(te + SR_PF * SR_TABW + 9)[0];
        if (MODE != SR_BUILD_X) {
          const double2 *q = rec + (size_t)(cw[kr] >> 32) * rstep;
          ra[kr] = MODE == SR_BUILD_N ? ld_plain2(q) : ld_stream2(q, pol);
          rb[kr] = MODE == SR_BUILD_N ? ld_plain2(q + n) : ld_stream2(q + n, pol);
        }
        te += SR_TABW;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + buf);  // this warp has read its last node of the band
    if (MODE == SR_DL || MODE == SR_SL) {
      // fixed order: a lane's points in stream order, then the streams -- one writer per target, no atomics.  Only the NS
      // warps of a target group meet here (named barrier 1 + grp); the buffer alternates with the item's parity.
      // streams 1 .. NS-1 hand their sums to stream 0 of their group: [item parity][group][stream - 1][3][tpw]
      const size_t gsz = (size_t)(NS - 1) * 3 * tpw;
      double *sg = s_red + ((size_t)(it & 1) * a.ngrp + grp) * gsz;
      if (strm > 0 && is_t) {
        double *sr = sg + (size_t)(strm - 1) * 3 * tpw + lane;
        sr[0] = pv0, sr[tpw] = pv1, sr[2 * tpw] = pv2;
      }
      asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(NS * 32) : "memory");
      if (strm == 0 && is_t) {
        const int tj = cell * a.npc + (jt0 + lane) * a.nlat + row;
        if (a.active[tj]) {
          const double cm = MODE == SR_DL ? a.coef * a.Bcell[cell] : a.coef;  // c2Mod, ModIntOnRbcs.F90:116
          double sum[3] = {pv0, pv1, pv2};
#pragma unroll
          for (int d = 0; d < 3; d++) {
            for (int q = 0; q < NS - 1; q++) sum[d] += sg[(size_t)q * 3 * tpw + d * tpw + lane];
            a.acc[(size_t)d * a.Np + tj] += cm * sum[d];
          }
        }
      }
    }
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


// ---------------------------------------------------------------------------------------------------------
static int planes_from_abi(rbc3d_ctx *c, const double *sp_abi, dbuf<double> &out) {
  Cells &C = c->cells;
  const size_t plane = (size_t)2 * C.nlat * C.nlon;
  RBC_TRY(out.resize((size_t)C.ncell * 12 * plane));
  const int *need = (c->prm.nranks > 1 && C.sg_cell_active.p) ? C.sg_cell_active.p : nullptr;
  k_spline_planes<<<c->sm_count * 8, 256, 0, c->stream>>>(C.ncell, 2 * C.nlat, C.nlon, need, sp_abi,
                                                           reinterpret_cast<double2 *>(out.p));
  KERNEL_CHECK();
  c->launches++;
  return RBC3D_OK;
}

template <int MODE>
static int launch_row(rbc3d_ctx *c, TargetList &t, const double *planes, double coef) {
  Cells &C = c->cells;
  if (C.sg_nactive == 0) return RBC3D_OK;
  RowArgs a;
  a.prm = c->prm;
  a.npc = C.npc, a.nlat = C.nlat, a.nlon = C.nlon, a.Np = C.Np;
  a.npts = C.sg_npatch_active, a.ntab = C.sg_ntab, a.ni_max = C.sg_ni_max, a.nslot = C.sg_nactive;
  a.ngrp = sr_groups(C.nlon), a.NS = sr_streams(C.nlon), a.tpw = sr_tpw(C.nlon);
  a.reps = std::max(1, std::min(c->sm_count / C.nlat, C.sg_nactive));
  a.nbuf = C.sg_K;
  a.tab = C.sg_st.p, a.rowinfo = C.sg_idx.p;
  a.planes = reinterpret_cast<const double2 *>(planes);
  a.cache = reinterpret_cast<double2 *>(C.sg_cache.p);
  a.spx = C.spx.p, a.th = C.th.p, a.phi = C.phi.p, a.Bcell = C.B.p;
  a.active = t.active.p, a.active_list = C.sg_active_list.p;
  a.tab_sl = c->tab_sl.p, a.tab_dl = c->tab_dl.p;
  a.coef = coef;
  a.acc = t.acc.p;
  CUDA_TRY(cudaFuncSetAttribute(k_sing_row<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C.sg_smem));
  k_sing_row<MODE><<<C.nlat * a.reps, a.ngrp * a.NS * 32 + 32, C.sg_smem, c->stream>>>(a);
  KERNEL_CHECK();
  c->launches++;
  return RBC3D_OK;
}

// geometry time (SourceList_UpdateCoord): density-independent cache of the patch integrand
int singular_prepare(rbc3d_ctx *c) {
  Cells &C = c->cells;
  C.sg_cache_ok = false;
  C.spFi_valid = false;
  if (!C.sg_ok || C.Np == 0 || c->sing_cache_mode == 0 || C.sg_nactive == 0) return RBC3D_OK;
  const size_t plane = (size_t)2 * C.nlat * C.nlon;
  const size_t per_cell = (size_t)C.nlat * C.sg_npatch_active * C.nlon;  // 32-byte records
  const size_t ncache = (size_t)C.sg_nactive;
  const size_t planes_bytes = (size_t)C.ncell * 12 * plane * sizeof(double);
  if (C.sg_cache.n < per_cell * ncache) {
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    // leave room for the two plane buffers (density spline; geometry / single-layer scratch) and the rest of the working set
    const size_t reserve = (C.spGi.n ? 0 : planes_bytes) + (C.spTi.n ? 0 : planes_bytes) + ((size_t)2 << 30);
    if (per_cell * ncache * sizeof(double4) + reserve > free_b + C.sg_cache.n * sizeof(double4)) return RBC3D_OK;  // direct kernel
    if (C.sg_cache.resize(per_cell * ncache) != RBC3D_OK) return RBC3D_OK;
  }
  TargetList &t = c->tl[RBC3D_TL_CELLS];
  RBC_TRY(planes_from_abi(c, C.spx.p, C.spTi));
  RBC_TRY(launch_row<SR_BUILD_X>(c, t, C.spTi.p, 0.0));
  RBC_TRY(planes_from_abi(c, C.spa3.p, C.spTi));
  RBC_TRY(launch_row<SR_BUILD_N>(c, t, C.spTi.p, 0.0));
  C.sg_cache_ok = true;
  if (C.g_set && !C.spGi_valid) RBC_TRY(singular_density_prepare(c));
  return RBC3D_OK;
}

// density time (SourceList_UpdateDensity with a host spline): node-interleaved copy of spline(g detJ)
int singular_density_prepare(rbc3d_ctx *c) {
  Cells &C = c->cells;
  if (!C.sg_cache_ok || !C.g_set || C.Np == 0 || C.spG.n == 0) return RBC3D_OK;
  RBC_TRY(planes_from_abi(c, C.spG.p, C.spGi));
  C.spGi_valid = true;
  return RBC3D_OK;
}

int singular_apply(rbc3d_ctx *c, TargetList &t, double c1, double c2) {
  if (t.kind != RBC3D_TL_CELLS) return RBC3D_OK;  // only on-surface targets (ModIntOnRbcs.F90:115)
  Cells &C = c->cells;
  if (C.Np == 0) return RBC3D_OK;
  bool sl = (c1 != 0), dl = (c2 != 0);
  if (!sl && !dl) return RBC3D_OK;
  if (C.sg_cache_ok) {
    if (dl && C.spGi_valid) {
      RBC_TRY(launch_row<SR_DL>(c, t, C.spGi.p, c2));
      dl = false;
    }
    if (sl && C.spF.n > 0) {
      if (!C.spFi_valid) {
        RBC_TRY(planes_from_abi(c, C.spF.p, C.spTi));
        C.spFi_valid = true;
      }
      RBC_TRY(launch_row<SR_SL>(c, t, C.spTi.p, c1));
      sl = false;
    }
    if (!sl && !dl) return RBC3D_OK;
  }
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
