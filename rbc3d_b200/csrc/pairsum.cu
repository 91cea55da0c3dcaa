// pairsum.cu -- real-space Ewald pair sum over the cell list (the pair loop of AddIntOnRbcs,
// ModIntOnRbcs.F90:47-112), the neighbour-cell scan that feeds the near-singular correction (:86-88,
// ModRbcSingInt.F90:341-364), the linear double-layer term (AddLinearInt, ModIntOnRbcs.F90:162-201) and the
// final division by the target coefficient.
//
// Layout: sources sorted by real-space cell (SoA, FP64); one warp per tile of <= 32 targets of one cell, lane =
// target, the 27 neighbour cells are walked with warp-uniform source loads, the range test is bit-exact
// (norm2_exact / min_image / rc2_thr) and the kernel evaluation is predicated per lane.
#include <algorithm>

#include "device_math.cuh"
#include "rbc3d_internal.h"

namespace rbc3d {

constexpr int PAIR_WARPS = 8;

struct PairArgs {
  Params prm;
  // sources (sorted by cell)
  const int *sstart;
  const double *sx;    // SoA(3,Np) sorted
  const double *sa3;   // sorted
  const double *sf;    // sorted
  const double *sgB;   // sorted, g * Bcoef
  const int *sorder;   // sorted position -> original source index
  int Np, npc, nlat, nlon;
  // targets
  const int2 *tiles;
  int ntiles;
  const int *tstart;   // offsets of the target cell list
  const int *torder;   // sorted position -> original target index
  const double *tx;    // SoA(3,nt) original order
  const int *tsurf;    // cell index of a cell target (or -1), original order
  int nt;
  int t_is_cells;      // targets are the cell points themselves (original index = source index)
  // tables
  const double *tab_sl;  // interleaved (c1,c2)
  const double *tab_dl;
  const double *omm;     // one-minus-mask [ilat_i][ilat_j][dl], dl = cyclic |dlon| in [0, nlon/2]
  const int *dlonmax;
  double c1, c2;
  double *acc;  // SoA(3,nt), += result
};

// Walk all sources in the 27 neighbour cells of cell c; calls body(j, pj, xx, yy, zz, r2, in_range) warp-uniformly
// for every source for which at least one lane is in range.  Sources are fetched 32 at a time with coalesced loads
// (lane = source) and handed round with shuffles (lane = target); sources of surface `excl_cell` (warp-uniform,
// -2 = none) are dropped before the distance test -- the same-surface pairs of cell targets belong to pairself.cu.
template <class Body>
__device__ __forceinline__ void walk_neighbours(const Params &prm, const int *__restrict__ sstart,
                                                const double *__restrict__ sx, const int *__restrict__ sorder, int Np,
                                                int npc, int c, bool valid, double xi, double yi, double zi,
                                                int excl_cell, Body body) {
  const int Nc1 = prm.Nc[0], Nc2 = prm.Nc[1], Nc3 = prm.Nc[2];
  const int c1 = c % Nc1, c2 = (c / Nc1) % Nc2, c3 = c / (Nc1 * Nc2);
  const double *__restrict__ sy = sx + Np;
  const double *__restrict__ sz = sx + 2 * (size_t)Np;
  const int lane = threadIdx.x & 31;
  for (int d3 = -1; d3 <= 1; d3++) {
    int n3 = c3 + d3;
    n3 = n3 < 0 ? n3 + Nc3 : (n3 >= Nc3 ? n3 - Nc3 : n3);
    for (int d2 = -1; d2 <= 1; d2++) {
      int n2 = c2 + d2;
      n2 = n2 < 0 ? n2 + Nc2 : (n2 >= Nc2 ? n2 - Nc2 : n2);
      for (int d1 = -1; d1 <= 1; d1++) {
        int n1 = c1 + d1;
        n1 = n1 < 0 ? n1 + Nc1 : (n1 >= Nc1 ? n1 - Nc1 : n1);
        const int nc = n1 + Nc1 * (n2 + Nc2 * n3);
        const int jb = sstart[nc], je = sstart[nc + 1];
        // sources of a cell are in ascending index order, so its surfaces are the range [first, last]: a cell that
        // holds only the excluded surface (the usual case in a dilute suspension) is skipped without loading it
        if (excl_cell >= 0 && je > jb && __ldg(sorder + jb) / npc == excl_cell && __ldg(sorder + je - 1) / npc == excl_cell)
          continue;
        for (int j0 = jb; j0 < je; j0 += 32) {
          const int jl = j0 + lane;
          const bool have = jl < je;
          double xs = 0, ys = 0, zs = 0;
          int pjl = -1;
          if (have) {
            xs = __ldg(sx + jl);
            ys = __ldg(sy + jl);
            zs = __ldg(sz + jl);
            pjl = __ldg(sorder + jl);
          }
          unsigned m = __ballot_sync(FULL_MASK, have && (pjl / npc) != excl_cell);
          while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const int pj = __shfl_sync(FULL_MASK, pjl, b);
            double xx = min_image(__dsub_rn(__shfl_sync(FULL_MASK, xs, b), xi), prm.iLb[0], prm.Lb[0]);
            double yy = min_image(__dsub_rn(__shfl_sync(FULL_MASK, ys, b), yi), prm.iLb[1], prm.Lb[1]);
            double zz = min_image(__dsub_rn(__shfl_sync(FULL_MASK, zs, b), zi), prm.iLb[2], prm.Lb[2]);
            double r2 = norm2_exact(xx, yy, zz);
            bool in = valid && !(r2 > prm.rc2_thr);
            if (__any_sync(FULL_MASK, in)) body(j0 + b, pj, xx, yy, zz, r2, in);
          }
        }
      }
    }
  }
}

template <bool SL, bool DL, bool EXCL>
__global__ void __launch_bounds__(PAIR_WARPS * 32) k_pair(PairArgs a) {
  extern __shared__ double smem[];
  // tables in shared memory: DL 8193, SL 2*8193
  double *s_dl = smem;
  double *s_sl = smem + (DL ? (RBC3D_NTAB + 1) : 0);
  for (int i = threadIdx.x; i <= RBC3D_NTAB; i += blockDim.x) {
    if (DL) s_dl[i] = a.tab_dl[i];
    if (SL) {
      s_sl[2 * i] = a.tab_sl[2 * i];
      s_sl[2 * i + 1] = a.tab_sl[2 * i + 1];
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // persistent CTAs: the lookup tables are staged once per CTA, warps stride over the target tiles
  for (int tile = blockIdx.x * PAIR_WARPS + warp; tile < a.ntiles; tile += gridDim.x * PAIR_WARPS) {
  const int2 tl = a.tiles[tile];
  const int c = tl.x;
  const int kend = min(tl.y + 32, a.tstart[c + 1]);
  const int k = tl.y + lane;
  const bool valid = k < kend;
  const int ti = valid ? a.torder[k] : 0;
  double xi = 0, yi = 0, zi = 0;
  int my_cell = -2, my_lat = 0, my_lon = 0;
  if (valid) {
    xi = a.tx[ti];
    yi = a.tx[(size_t)a.nt + ti];
    zi = a.tx[2 * (size_t)a.nt + ti];
    my_cell = a.tsurf[ti];
    if (my_cell >= 0) {
      int rem = ti - my_cell * a.npc;  // cell targets: original index = cell*npc + ilon*nlat + ilat
      my_lon = rem / a.nlat;
      my_lat = rem - my_lon * a.nlat;
    }
  }
  double ax = 0, ay = 0, az = 0;    // SL sum
  double bx = 0, by = 0, bz = 0;    // DL sum
  const double r_eps2 = a.prm.r_eps * a.prm.r_eps;
  const int nlonh = a.nlon / 2 + 1;
  const size_t Np = a.Np;

  // same-surface sources are excluded (EXCL: pairself.cu owns them): warp-uniformly when the whole tile belongs to
  // one surface, per lane otherwise
  int excl = -2;
  if (EXCL) {
    const int c0 = __shfl_sync(FULL_MASK, my_cell, __ffs(__ballot_sync(FULL_MASK, valid)) - 1);
    if (__all_sync(FULL_MASK, !valid || my_cell == c0)) excl = c0;
  }
  walk_neighbours(a.prm, a.sstart, a.sx, a.sorder, a.Np, a.npc, c, valid, xi, yi, zi, excl,
                  [&](int j, int pj, double xx, double yy, double zz, double r2, bool in) {
    const int cj = pj / a.npc;
    if (in && r2 >= r_eps2 && !(EXCL && cj == my_cell)) {
      double om = 1.0;  // 1 - mask
      if (!EXCL && cj == my_cell) {
        int rem = pj - cj * a.npc;
        int lon_j = rem / a.nlat, lat_j = rem - lon_j * a.nlat;
        int dl = abs(my_lon - lon_j);
        dl = min(dl, a.nlon - dl);
        int pair = my_lat * a.nlat + lat_j;
        if (dl <= __ldg(a.dlonmax + pair)) om = __ldg(a.omm + (size_t)pair * nlonh + dl);
      }
      const double rinv = rsqrt_pos(r2);
      const double r = r2 * rinv;
      const double s = r * a.prm.tab_scale;
      const int i = (int)s;
      if (i < RBC3D_NTAB) {
        const double fr = s - (double)i;
        if (SL) {
          const double t10 = s_sl[2 * i], t20 = s_sl[2 * i + 1], t11 = s_sl[2 * i + 2], t21 = s_sl[2 * i + 3];
          const double e1 = fma(fr, t11 - t10, t10), e2 = fma(fr, t21 - t20, t20);
          const double ir2 = rinv * rinv;
          const double EA = e1 * rinv * ir2 + e2 * ir2;  // ModEwaldFunc.F90:126-127
          const double EB = e1 * rinv - e2;
          const double fx = __ldg(a.sf + j), fy = __ldg(a.sf + Np + j), fz = __ldg(a.sf + 2 * Np + j);
          const double xf = EA * (xx * fx + yy * fy + zz * fz);
          ax += om * (xf * xx + EB * fx);
          ay += om * (xf * yy + EB * fy);
          az += om * (xf * zz + EB * fz);
        }
        if (DL) {
          const double t0 = s_dl[i], t1 = s_dl[i + 1];
          const double e = fma(fr, t1 - t0, t0);
          const double ir2 = rinv * rinv;
          const double EA = e * ir2 * ir2 * rinv;  // c1 / r^5, ModEwaldFunc.F90:174
          const double gx = __ldg(a.sgB + j), gy = __ldg(a.sgB + Np + j), gz = __ldg(a.sgB + 2 * Np + j);
          const double nx = __ldg(a.sa3 + j), ny = __ldg(a.sa3 + Np + j), nz = __ldg(a.sa3 + 2 * Np + j);
          const double q = om * EA * (xx * gx + yy * gy + zz * gz) * (xx * nx + yy * ny + zz * nz);
          bx += q * xx;
          by += q * yy;
          bz += q * zz;
        }
      }
    }
  });

  if (valid) {
    double *acc = a.acc;
    acc[ti] += a.c1 * ax + a.c2 * bx;
    acc[(size_t)a.nt + ti] += a.c1 * ay + a.c2 * by;
    acc[2 * (size_t)a.nt + ti] += a.c1 * az + a.c2 * bz;
  }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Per-geometry pair lists.  Between cells of different surfaces (and for off-surface targets) only a small part of
// the sources in the 27 neighbour cells is within rc of ANY target of a tile.  The geometry is fixed across the GMRES
// matvecs of a time step, so the walk is done once per geometry (k_pair_scan: count, then fill, in walk order) and
// every operator application only visits the listed (tile, source) entries (k_pair_list).  Same arithmetic and
// summation order as k_pair.
template <bool FILL, bool EXCL>
__global__ void __launch_bounds__(PAIR_WARPS * 32) k_pair_scan(PairArgs a, int *__restrict__ cnt,
                                                                const int *__restrict__ off, int *__restrict__ src) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * PAIR_WARPS + warp;
  if (tile >= a.ntiles) return;
  const int2 tl = a.tiles[tile];
  const int c = tl.x;
  const int kend = min(tl.y + 32, a.tstart[c + 1]);
  const int k = tl.y + lane;
  const bool valid = k < kend;
  const int ti = valid ? a.torder[k] : 0;
  double xi = 0, yi = 0, zi = 0;
  int my_cell = -2;
  if (valid) {
    xi = a.tx[ti];
    yi = a.tx[(size_t)a.nt + ti];
    zi = a.tx[2 * (size_t)a.nt + ti];
    my_cell = a.tsurf[ti];
  }
  int excl = -2;
  if (EXCL) {
    const int c0 = __shfl_sync(FULL_MASK, my_cell, __ffs(__ballot_sync(FULL_MASK, valid)) - 1);
    if (__all_sync(FULL_MASK, !valid || my_cell == c0)) excl = c0;
  }
  int n = 0;
  const int o = FILL ? off[tile] : 0;
  walk_neighbours(a.prm, a.sstart, a.sx, a.sorder, a.Np, a.npc, c, valid, xi, yi, zi, excl,
                  [&](int j, int pj, double, double, double, double, bool in) {
                    const int cj = pj / a.npc;
                    if (__any_sync(FULL_MASK, in && !(EXCL && cj == my_cell))) {
                      if (FILL && lane == 0) src[o + n] = j;
                      n++;
                    }
                  });
  if (!FILL && lane == 0) cnt[tile] = n;
}

template <bool SL, bool DL, bool EXCL>
__global__ void __launch_bounds__(PAIR_WARPS * 32) k_pair_list(PairArgs a, const int *__restrict__ off,
                                                                const int *__restrict__ src) {
  extern __shared__ double smem[];
  double *s_dl = smem;
  double *s_sl = smem + (DL ? (RBC3D_NTAB + 1) : 0);
  for (int i = threadIdx.x; i <= RBC3D_NTAB; i += blockDim.x) {
    if (DL) s_dl[i] = a.tab_dl[i];
    if (SL) {
      s_sl[2 * i] = a.tab_sl[2 * i];
      s_sl[2 * i + 1] = a.tab_sl[2 * i + 1];
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double r_eps2 = a.prm.r_eps * a.prm.r_eps;
  const int nlonh = a.nlon / 2 + 1;
  const size_t Np = a.Np;
  const double *__restrict__ sy = a.sx + Np;
  const double *__restrict__ sz = a.sx + 2 * Np;
  for (int tile = blockIdx.x * PAIR_WARPS + warp; tile < a.ntiles; tile += gridDim.x * PAIR_WARPS) {
    const int eb = off[tile], ee = off[tile + 1];
    if (eb == ee) continue;
    const int2 tl = a.tiles[tile];
    const int c = tl.x;
    const int kend = min(tl.y + 32, a.tstart[c + 1]);
    const int k = tl.y + lane;
    const bool valid = k < kend;
    const int ti = valid ? a.torder[k] : 0;
    double xi = 0, yi = 0, zi = 0;
    int my_cell = -2, my_lat = 0, my_lon = 0;
    if (valid) {
      xi = a.tx[ti];
      yi = a.tx[(size_t)a.nt + ti];
      zi = a.tx[2 * (size_t)a.nt + ti];
      my_cell = a.tsurf[ti];
      if (my_cell >= 0) {
        int rem = ti - my_cell * a.npc;
        my_lon = rem / a.nlat;
        my_lat = rem - my_lon * a.nlat;
      }
    }
    double ax = 0, ay = 0, az = 0, bx = 0, by = 0, bz = 0;
    for (int e0 = eb; e0 < ee; e0 += 32) {
      // 32 listed sources at a time: lane = source for the loads, then handed round with shuffles
      const int el = e0 + lane;
      const bool have = el < ee;
      int jl = 0, pjl = -1;
      double xs = 0, ys = 0, zs = 0;
      if (have) {
        jl = __ldg(src + el);
        xs = __ldg(a.sx + jl);
        ys = __ldg(sy + jl);
        zs = __ldg(sz + jl);
        pjl = __ldg(a.sorder + jl);
      }
      const int nb = min(32, ee - e0);
      for (int b = 0; b < nb; b++) {
        const int j = __shfl_sync(FULL_MASK, jl, b);
        const int pj = __shfl_sync(FULL_MASK, pjl, b);
        const double xx = min_image(__dsub_rn(__shfl_sync(FULL_MASK, xs, b), xi), a.prm.iLb[0], a.prm.Lb[0]);
        const double yy = min_image(__dsub_rn(__shfl_sync(FULL_MASK, ys, b), yi), a.prm.iLb[1], a.prm.Lb[1]);
        const double zz = min_image(__dsub_rn(__shfl_sync(FULL_MASK, zs, b), zi), a.prm.iLb[2], a.prm.Lb[2]);
        const double r2 = norm2_exact(xx, yy, zz);
        const bool in = valid && !(r2 > a.prm.rc2_thr);
        const int cj = pj / a.npc;
        if (in && r2 >= r_eps2 && !(EXCL && cj == my_cell)) {
          double om = 1.0;  // 1 - mask
          if (!EXCL && cj == my_cell) {
            int rem = pj - cj * a.npc;
            int lon_j = rem / a.nlat, lat_j = rem - lon_j * a.nlat;
            int dl = abs(my_lon - lon_j);
            dl = min(dl, a.nlon - dl);
            int pair = my_lat * a.nlat + lat_j;
            if (dl <= __ldg(a.dlonmax + pair)) om = __ldg(a.omm + (size_t)pair * nlonh + dl);
          }
          const double rinv = rsqrt_pos(r2);
          const double r = r2 * rinv;
          const double s = r * a.prm.tab_scale;
          const int i = (int)s;
          if (i < RBC3D_NTAB) {
            const double fr = s - (double)i;
            if (SL) {
              const double t10 = s_sl[2 * i], t20 = s_sl[2 * i + 1], t11 = s_sl[2 * i + 2], t21 = s_sl[2 * i + 3];
              const double e1 = fma(fr, t11 - t10, t10), e2 = fma(fr, t21 - t20, t20);
              const double ir2 = rinv * rinv;
              const double EA = e1 * rinv * ir2 + e2 * ir2;  // ModEwaldFunc.F90:126-127
              const double EB = e1 * rinv - e2;
              const double fx = __ldg(a.sf + j), fy = __ldg(a.sf + Np + j), fz = __ldg(a.sf + 2 * Np + j);
              const double xf = EA * (xx * fx + yy * fy + zz * fz);
              ax += om * (xf * xx + EB * fx);
              ay += om * (xf * yy + EB * fy);
              az += om * (xf * zz + EB * fz);
            }
            if (DL) {
              const double t0 = s_dl[i], t1 = s_dl[i + 1];
              const double e = fma(fr, t1 - t0, t0);
              const double ir2 = rinv * rinv;
              const double EA = e * ir2 * ir2 * rinv;  // c1 / r^5, ModEwaldFunc.F90:174
              const double gx = __ldg(a.sgB + j), gy = __ldg(a.sgB + Np + j), gz = __ldg(a.sgB + 2 * Np + j);
              const double nx = __ldg(a.sa3 + j), ny = __ldg(a.sa3 + Np + j), nz = __ldg(a.sa3 + 2 * Np + j);
              const double q = om * EA * (xx * gx + yy * gy + zz * gz) * (xx * nx + yy * ny + zz * nz);
              bx += q * xx;
              by += q * yy;
              bz += q * zz;
            }
          }
        }
      }
    }
    if (valid) {
      double *acc = a.acc;
      acc[ti] += a.c1 * ax + a.c2 * bx;
      acc[(size_t)a.nt + ti] += a.c1 * ay + a.c2 * by;
      acc[2 * (size_t)a.nt + ti] += a.c1 * az + a.c2 * bz;
    }
  }
}

static void fill_args(rbc3d_ctx *c, TargetList &t, PairArgs &a) {
  Cells &C = c->cells;
  a.prm = c->prm;
  a.sstart = C.cl.start.p;
  a.sx = C.sx.p;
  a.sa3 = C.sa3.p;
  a.sf = C.sf.p;
  a.sgB = C.sgB.p;
  a.sorder = C.cl.order.p;
  a.Np = C.Np;
  a.npc = C.npc;
  a.nlat = C.nlat;
  a.nlon = C.nlon;
  a.tiles = t.tiles.p;
  a.ntiles = t.ntiles;
  const bool tc = (t.kind == RBC3D_TL_CELLS);
  CellList &tcl = t.cl;
  a.tstart = tcl.start.p;
  a.torder = tcl.order.p;
  a.tx = t.x.p;
  a.tsurf = t.surf.p;
  a.nt = t.n;
  a.t_is_cells = tc;
  a.tab_sl = c->tab_sl.p;
  a.tab_dl = c->tab_dl.p;
  a.omm = C.omm.p;
  a.dlonmax = C.dlonmax.p;
  a.acc = t.acc.p;
}

int pair_sum(rbc3d_ctx *c, TargetList &t, double c1, double c2) {
  if (t.ntiles == 0 || c->cells.Np == 0) return RBC3D_OK;
  PairArgs a;
  fill_args(c, t, a);
  a.c1 = c1;
  a.c2 = c2;
  const bool sl = (c1 != 0), dl = (c2 != 0);  // ModIntOnRbcs.F90:91,99 test c /= 0 exactly
  if (!sl && !dl) return RBC3D_OK;
  const size_t sm = sizeof(double) * (RBC3D_NTAB + 1) * ((dl ? 1 : 0) + (sl ? 2 : 0));
  // persistent grid: as many CTAs as fit (shared memory bound), a few waves of tiles each
  const int per_sm = std::max(1, std::min(8, (int)((size_t)(227 * 1024) / (sm + 1024))));
  const int grid = std::min((t.ntiles + PAIR_WARPS - 1) / PAIR_WARPS, c->sm_count * per_sm);
  const bool excl = pairself_available(c, t);  // same-surface pairs: dense per-cell kernel (pairself.cu)
  static const bool no_list = getenv("RBC3D_PAIR_NO_LIST") != nullptr;
  const bool use_list = !no_list && (excl || t.kind != RBC3D_TL_CELLS);
  if (use_list) {
    // (tile, source) entries with a source within rc of a target of the tile: once per geometry
    if (!t.plist_valid || t.plist_excl != (int)excl || t.plist_geom != c->cells.geom_version) {
      const int sgrid = (t.ntiles + PAIR_WARPS - 1) / PAIR_WARPS;
      RBC_TRY(t.plist_off.resize((size_t)t.ntiles + 2));
      CUDA_TRY(cudaMemsetAsync(t.plist_off.p, 0, sizeof(int) * ((size_t)t.ntiles + 2), c->stream));
      if (excl)
        k_pair_scan<false, true><<<sgrid, PAIR_WARPS * 32, 0, c->stream>>>(a, t.plist_off.p, nullptr, nullptr);
      else
        k_pair_scan<false, false><<<sgrid, PAIR_WARPS * 32, 0, c->stream>>>(a, t.plist_off.p, nullptr, nullptr);
      KERNEL_CHECK();
      RBC_TRY(device_exclusive_scan(c, t.plist_off.p, t.ntiles + 1));
      int nent = 0;
      CUDA_TRY(cudaMemcpyAsync(&nent, t.plist_off.p + t.ntiles, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      CUDA_TRY(cudaStreamSynchronize(c->stream));
      RBC_TRY(t.plist_src.resize(nent > 0 ? nent : 1));
      if (nent > 0) {
        if (excl)
          k_pair_scan<true, true><<<sgrid, PAIR_WARPS * 32, 0, c->stream>>>(a, nullptr, t.plist_off.p, t.plist_src.p);
        else
          k_pair_scan<true, false><<<sgrid, PAIR_WARPS * 32, 0, c->stream>>>(a, nullptr, t.plist_off.p, t.plist_src.p);
        KERNEL_CHECK();
      }
      c->launches += 3;
      t.plist_n = nent;
      t.plist_valid = true;
      t.plist_excl = (int)excl;
      t.plist_geom = c->cells.geom_version;
    }
    if (t.plist_n > 0) {
#define LAUNCH_LIST(SL_, DL_, EX_)                                                                                  \
  do {                                                                                                              \
    CUDA_TRY(cudaFuncSetAttribute(k_pair_list<SL_, DL_, EX_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    k_pair_list<SL_, DL_, EX_><<<grid, PAIR_WARPS * 32, sm, c->stream>>>(a, t.plist_off.p, t.plist_src.p);         \
  } while (0)
      if (sl && dl) {
        if (excl) LAUNCH_LIST(true, true, true); else LAUNCH_LIST(true, true, false);
      } else if (sl) {
        if (excl) LAUNCH_LIST(true, false, true); else LAUNCH_LIST(true, false, false);
      } else {
        if (excl) LAUNCH_LIST(false, true, true); else LAUNCH_LIST(false, true, false);
      }
#undef LAUNCH_LIST
      KERNEL_CHECK();
      c->launches++;
    }
  } else {
#define LAUNCH_PAIR(SL_, DL_, EX_)                                                                              \
  do {                                                                                                          \
    CUDA_TRY(cudaFuncSetAttribute(k_pair<SL_, DL_, EX_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    k_pair<SL_, DL_, EX_><<<grid, PAIR_WARPS * 32, sm, c->stream>>>(a);                                        \
  } while (0)
    if (sl && dl) {
      if (excl) LAUNCH_PAIR(true, true, true); else LAUNCH_PAIR(true, true, false);
    } else if (sl) {
      if (excl) LAUNCH_PAIR(true, false, true); else LAUNCH_PAIR(true, false, false);
    } else {
      if (excl) LAUNCH_PAIR(false, true, true); else LAUNCH_PAIR(false, true, false);
    }
#undef LAUNCH_PAIR
    KERNEL_CHECK();
    c->launches++;
  }
  if (excl) RBC_TRY(pairself_apply(c, t, c1, c2));
  return RBC3D_OK;
}

// ---- neighbour set signature (bit-exactness test of the cell list + range test) ----
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}

__global__ void __launch_bounds__(PAIR_WARPS * 32) k_signature(PairArgs a, int *__restrict__ count,
                                                               unsigned long long *__restrict__ sig) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * PAIR_WARPS + warp;
  if (tile >= a.ntiles) return;
  const int2 tl = a.tiles[tile];
  const int c = tl.x;
  const int kend = min(tl.y + 32, a.tstart[c + 1]);
  const int k = tl.y + lane;
  const bool valid = k < kend;
  const int ti = valid ? a.torder[k] : 0;
  double xi = 0, yi = 0, zi = 0;
  if (valid) {
    xi = a.tx[ti];
    yi = a.tx[(size_t)a.nt + ti];
    zi = a.tx[2 * (size_t)a.nt + ti];
  }
  int cnt = 0;
  unsigned long long s = 0;
  walk_neighbours(a.prm, a.sstart, a.sx, a.sorder, a.Np, a.npc, c, valid, xi, yi, zi, -2,
                  [&](int, int pj, double, double, double, double, bool in) {
                    if (in) {
                      cnt++;
                      s += mix64((unsigned long long)pj);
                    }
                  });
  if (valid) {
    count[ti] = cnt;
    sig[ti] = s;
  }
}

int neighbor_signature(rbc3d_ctx *c, TargetList &t, int *count, unsigned long long *sig) {
  PairArgs a;
  fill_args(c, t, a);
  a.c1 = a.c2 = 0;
  dbuf<int> dc;
  dbuf<unsigned long long> ds;
  RBC_TRY(dc.resize(t.n));
  RBC_TRY(ds.resize(t.n));
  CUDA_TRY(cudaMemsetAsync(dc.p, 0, sizeof(int) * t.n, c->stream));
  CUDA_TRY(cudaMemsetAsync(ds.p, 0, sizeof(unsigned long long) * t.n, c->stream));
  if (t.ntiles > 0) {
    k_signature<<<(t.ntiles + PAIR_WARPS - 1) / PAIR_WARPS, PAIR_WARPS * 32, 0, c->stream>>>(a, dc.p, ds.p);
    KERNEL_CHECK();
  }
  CUDA_TRY(cudaMemcpyAsync(count, dc.p, sizeof(int) * t.n, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaMemcpyAsync(sig, ds.p, sizeof(unsigned long long) * t.n, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  dc.release();
  ds.release();
  return RBC3D_OK;
}

// ---- neighbour-cell scan: for every target the other cells with a source within rc and the closest
//      mesh point of each (NbrRbcList_Insert, ModRbcSingInt.F90:341-364).  Two passes: count, then fill. ----
template <bool FILL>
__global__ void __launch_bounds__(PAIR_WARPS * 32) k_nbr_scan(PairArgs a, int *__restrict__ cnt,
                                                               const int *__restrict__ off, int *__restrict__ e_target,
                                                               int *__restrict__ e_cell, int *__restrict__ e_pt,
                                                               int *__restrict__ overflow) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * PAIR_WARPS + warp;
  if (tile >= a.ntiles) return;
  const int2 tl = a.tiles[tile];
  const int c = tl.x;
  const int kend = min(tl.y + 32, a.tstart[c + 1]);
  const int k = tl.y + lane;
  const bool valid = k < kend;
  const int ti = valid ? a.torder[k] : 0;
  double xi = 0, yi = 0, zi = 0;
  int my_cell = -2;
  if (valid) {
    xi = a.tx[ti];
    yi = a.tx[(size_t)a.nt + ti];
    zi = a.tx[2 * (size_t)a.nt + ti];
    my_cell = a.tsurf[ti];
  }
  int n = 0;
  int l_cell[RBC3D_NBR_MAX], l_pt[RBC3D_NBR_MAX];
  double l_r2[RBC3D_NBR_MAX];
  bool ovf = false;
  int excl = -2;  // cell targets: the own surface never enters the neighbour list
  {
    const int c0 = __shfl_sync(FULL_MASK, my_cell, __ffs(__ballot_sync(FULL_MASK, valid)) - 1);
    if (c0 >= 0 && __all_sync(FULL_MASK, !valid || my_cell == c0)) excl = c0;
  }
  walk_neighbours(a.prm, a.sstart, a.sx, a.sorder, a.Np, a.npc, c, valid, xi, yi, zi, excl,
                  [&](int, int pj, double, double, double, double r2, bool in) {
                    const int cj = pj / a.npc;
                    if (in && cj != my_cell) {
                      int q = 0;
                      for (; q < n; q++)
                        if (l_cell[q] == cj) break;
                      if (q < n) {
                        // strict "<" on rr (ModRbcSingInt.F90:352); sqrt is monotone, ties keep the first
                        if (r2 < l_r2[q]) {
                          l_r2[q] = r2;
                          l_pt[q] = pj - cj * a.npc;
                        }
                      } else if (n < RBC3D_NBR_MAX) {
                        l_cell[n] = cj;
                        l_r2[n] = r2;
                        l_pt[n] = pj - cj * a.npc;
                        n++;
                      } else {
                        ovf = true;
                      }
                    }
                  });
  if (!valid) return;
  if (ovf) atomicExch(overflow, 1);
  if (!FILL) {
    cnt[k] = n;
  } else {
    int o = off[k];
    for (int q = 0; q < n; q++) {
      e_target[o + q] = ti;
      e_cell[o + q] = l_cell[q];
      e_pt[o + q] = l_pt[q];
    }
  }
}

int nearsing_scan(rbc3d_ctx *c, TargetList &t, bool fill) {
  PairArgs a;
  fill_args(c, t, a);
  a.c1 = a.c2 = 0;
  NearSing &ns = t.ns;
  const int grid = (t.ntiles + PAIR_WARPS - 1) / PAIR_WARPS;
  if (grid == 0) return RBC3D_OK;
  if (!fill)
    k_nbr_scan<false><<<grid, PAIR_WARPS * 32, 0, c->stream>>>(a, ns.cnt.p, nullptr, nullptr, nullptr, nullptr,
                                                               ns.overflow.p);
  else
    k_nbr_scan<true><<<grid, PAIR_WARPS * 32, 0, c->stream>>>(a, nullptr, ns.off.p, ns.target.p, ns.cell.p,
                                                              ns.pt.p, ns.overflow.p);
  KERNEL_CHECK();
  c->launches++;
  return RBC3D_OK;
}

// ---- gather the sorted source copies ----
__global__ void k_gather3(int n, const int *__restrict__ order, const double *__restrict__ src,
                          double *__restrict__ dst, const double *__restrict__ scale_cell, int npc) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int p = order[k];
  double s = scale_cell ? scale_cell[p / npc] : 1.0;
  dst[k] = src[p] * s;
  dst[(size_t)n + k] = src[(size_t)n + p] * s;
  dst[2 * (size_t)n + k] = src[2 * (size_t)n + p] * s;
}

int cells_gather_sorted(rbc3d_ctx *c, bool geom, bool f, bool g) {
  Cells &C = c->cells;
  const int n = C.Np;
  if (n == 0) return RBC3D_OK;
  const int grid = (n + 255) / 256;
  if (geom) {
    RBC_TRY(C.sx.resize(3 * (size_t)n));
    RBC_TRY(C.sa3.resize(3 * (size_t)n));
    k_gather3<<<grid, 256, 0, c->stream>>>(n, C.cl.order.p, C.x.p, C.sx.p, nullptr, C.npc);
    k_gather3<<<grid, 256, 0, c->stream>>>(n, C.cl.order.p, C.a3.p, C.sa3.p, nullptr, C.npc);
    c->launches += 2;
  }
  if (f) {
    RBC_TRY(C.sf.resize(3 * (size_t)n));
    k_gather3<<<grid, 256, 0, c->stream>>>(n, C.cl.order.p, C.f.p, C.sf.p, nullptr, C.npc);
    c->launches++;
  }
  if (g) {
    RBC_TRY(C.sgB.resize(3 * (size_t)n));
    k_gather3<<<grid, 256, 0, c->stream>>>(n, C.cl.order.p, C.g.p, C.sgB.p, C.B.p, C.npc);
    c->launches++;
  }
  KERNEL_CHECK();
  return RBC3D_OK;
}

// ---- AddLinearInt (ModIntOnRbcs.F90:162-201): xvint = -8 pi / V * sum_j B_j x_j (g_j . a3_j) with g already
//      weighted by detJ*w; deterministic two-stage reduction ----
constexpr int LIN_BLOCKS = 296;
__global__ void __launch_bounds__(256) k_linear_partial(int n, int npc, int p_lo, int p_hi, const double *__restrict__ x,
                                                        const double *__restrict__ g, const double *__restrict__ a3,
                                                        const double *__restrict__ B, double *__restrict__ part) {
  double sx = 0, sy = 0, sz = 0;
  for (int p = p_lo + blockIdx.x * blockDim.x + threadIdx.x; p < p_hi; p += gridDim.x * blockDim.x) {
    double vn = g[p] * a3[p] + g[(size_t)n + p] * a3[(size_t)n + p] + g[2 * (size_t)n + p] * a3[2 * (size_t)n + p];
    vn *= B[p / npc];
    sx += x[p] * vn;
    sy += x[(size_t)n + p] * vn;
    sz += x[2 * (size_t)n + p] * vn;
  }
  __shared__ double sh[3][8];
  sx = warp_sum(sx);
  sy = warp_sum(sy);
  sz = warp_sum(sz);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sh[0][warp] = sx;
    sh[1][warp] = sy;
    sh[2][warp] = sz;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0;
    for (int w = 0; w < 8; w++) s += sh[threadIdx.x][w];
    part[threadIdx.x * LIN_BLOCKS + blockIdx.x] = s;
  }
}

__global__ void k_linear_final(const double *__restrict__ part, double scale, double *__restrict__ out) {
  if (threadIdx.x < 3) {
    double s = 0;
    for (int b = 0; b < LIN_BLOCKS; b++) s += part[threadIdx.x * LIN_BLOCKS + b];
    out[threadIdx.x] = s * scale;
  }
}

int linear_term(rbc3d_ctx *c, TargetList &t, double c2) {
  Cells &C = c->cells;
  RBC_TRY(C.xvint_part.resize(3 * LIN_BLOCKS + 8));
  if (c2 == 0 || C.Np == 0) {
    CUDA_TRY(cudaMemsetAsync(C.xvint_part.p + 3 * LIN_BLOCKS, 0, 3 * sizeof(double), c->stream));
    return RBC3D_OK;
  }
  const Params &p = c->prm;
  // several ranks: every rank sums a contiguous block of cells (a partition whatever the target ownership is) and the
  // three numbers are all-reduced, instead of every rank reading all N points (collective 5)
  const int c_lo = p.nranks > 1 ? (int)((long long)C.ncell * p.rank / p.nranks) : 0;
  const int c_hi = p.nranks > 1 ? (int)((long long)C.ncell * (p.rank + 1) / p.nranks) : C.ncell;
  k_linear_partial<<<LIN_BLOCKS, 256, 0, c->stream>>>(C.Np, C.npc, c_lo * C.npc, c_hi * C.npc, C.x.p, C.g.p, C.a3.p, C.B.p,
                                                      C.xvint_part.p);
  const double scale = -8.0 * RBC_PI * (p.iLb[0] * p.iLb[1] * p.iLb[2]) * c2;
  k_linear_final<<<1, 32, 0, c->stream>>>(C.xvint_part.p, scale, C.xvint_part.p + 3 * LIN_BLOCKS);
  KERNEL_CHECK();
  c->launches += 2;
  if (p.nranks > 1) RBC_TRY(comm_allreduce_sum(c, C.xvint_part.p + 3 * LIN_BLOCKS, 3));
  return RBC3D_OK;
}

// ---- v (+)= (acc + c2*xvint) / Acoef for active targets (the "/tlist%Acoef(i)" of every term) ----
__global__ void k_combine(int n, const double *__restrict__ acc, const double *__restrict__ acc2,
                          const double *__restrict__ xv, const double *__restrict__ A, const int *__restrict__ active,
                          double *__restrict__ v, int accumulate) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool on = active[i] != 0;
  const double ia = 1.0 / A[i];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    double s = acc[(size_t)d * n + i];
    if (acc2) s += acc2[(size_t)d * n + i];
    double r = on ? (s + xv[d]) * ia : 0.0;
    if (accumulate)
      v[(size_t)d * n + i] += r;
    else
      v[(size_t)d * n + i] = r;
  }
}

int combine(rbc3d_ctx *c, TargetList &t, double *v_dev, bool accumulate, const double *acc2) {
  if (t.n == 0) return RBC3D_OK;
  Cells &C = c->cells;
  k_combine<<<(t.n + 255) / 256, 256, 0, c->stream>>>(t.n, t.acc.p, acc2, C.xvint_part.p + 3 * LIN_BLOCKS, t.Acoef.p,
                                                      t.active.p, v_dev, accumulate ? 1 : 0);
  KERNEL_CHECK();
  c->launches++;
  return RBC3D_OK;
}

}  // namespace rbc3d
