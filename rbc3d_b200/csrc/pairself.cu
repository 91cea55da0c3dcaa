// pairself.cu -- the same-surface part of the real-space pair loop of AddIntOnRbcs (ModIntOnRbcs.F90:63-112,
// branch surfId_i == surfId_j): every target of a cell against every mesh point of the SAME cell, weighted by
// 1 - MaskFunc(DistOnSphere / patch radius).
//
// At the reference's cut-off (rc ~ 1.2 cell radii) ~43 % of a cell's own 2592 points are within range of each of its
// points, and in a suspension at physiological spacing those same-cell pairs are > 99 % of all in-range pairs.  The
// hashed cell list is the wrong tool for them (it returns the whole cell plus its neighbours as candidates), so they
// are evaluated as a dense block per cell:
//   * one CTA per (cell, third of its targets); the cell's sources (x, g*B, a3 or x, f) are staged chunk-wise in shared
//     memory as 80-byte records and read with warp-wide broadcast LDS.128;
//   * warp = a compact 4 lat x 8 lon patch of targets, so whole sources out of range of the patch are skipped
//     warp-uniformly; lane = one target, accumulators in registers;
//   * Ewald lookup table(s) in shared memory (lerp on 8192 intervals like EwaldCoeff_DL/SL, ModEwaldFunc.F90:86-178);
//   * the mask is non-zero for ~9 % of the pairs only: a 64-bit set per (lat_i, |dlon|) says for which lat_j the
//     one-minus-mask table has to be consulted;
//   * the range test uses the reference's arithmetic (no FMA contraction); the minimum-image step is skipped for
//     cells whose extent is below half a box (nint(xx/L) = 0 exactly), which is decided per cell at geometry time.
// Cross-surface pairs stay with the cell-list kernel (pairsum.cu), which now skips same-surface sources.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "device_math.cuh"
#include "rbc3d_internal.h"

namespace rbc3d {

constexpr int PS_WARPS = 27;          // 864 targets per CTA
constexpr int PS_REC = 10;            // doubles per staged source record (80 B)
constexpr int PS_CHUNK_MAX = 1296;    // sources per staged chunk

struct SelfArgs {
  Params prm;
  int ncell, npc, nlat, nlon, Np;
  const double *x, *a3, *f, *g;   // SoA(3,Np), original order; f, g weighted by detJ*w
  const double *Bcell;
  const int *warp_tgt;            // [nwarps_cell][32] mesh point of (warp, lane) or -1
  int nwarps_cell, ctas_per_cell;
  const unsigned long long *maskbits;  // [nlat][nlon/2+1]
  const double *omm;                    // [nlat][nlat][nlon/2+1]
  const double *tab_sl, *tab_dl;
  const int *active, *cell_active;
  const unsigned char *cell_compact;
  double c1, c2;
  double *acc;
  int chunk_cols;                 // longitude columns per staged chunk
};

template <bool SL>
__global__ void __launch_bounds__(PS_WARPS * 32, 1) k_pair_self(SelfArgs a) {
  extern __shared__ double smem[];
  // [table: DL 8193 | SL 2*8193][maskbits nlat*nlonh (as double-sized words)][records chunk*10]
  const int nlonh = a.nlon / 2 + 1;
  double *s_tab = smem;
  const int ntab = SL ? 2 * (RBC3D_NTAB + 1) : (RBC3D_NTAB + 1) + 1;   // keep 16-byte alignment of what follows
  unsigned long long *s_bits = reinterpret_cast<unsigned long long *>(smem + ntab);
  const int nbits = (a.nlat * nlonh + 1) & ~1;
  double *s_rec = smem + ntab + nbits;
  const int cell = blockIdx.x / a.ctas_per_cell, part = blockIdx.x - cell * a.ctas_per_cell;
  if (!a.cell_active[cell]) return;
  for (int i = threadIdx.x; i <= RBC3D_NTAB; i += blockDim.x) {
    if (SL) {
      s_tab[2 * i] = a.tab_sl[2 * i];
      s_tab[2 * i + 1] = a.tab_sl[2 * i + 1];
    } else {
      s_tab[i] = a.tab_dl[i];
    }
  }
  for (int i = threadIdx.x; i < a.nlat * nlonh; i += blockDim.x) s_bits[i] = a.maskbits[i];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wcell = part * PS_WARPS + warp;  // warp index inside the cell
  int p = -1;
  if (wcell < a.nwarps_cell) p = a.warp_tgt[wcell * 32 + lane];
  const size_t base = (size_t)cell * a.npc, Np = a.Np;
  bool valid = p >= 0;
  if (valid) valid = a.active[base + p] != 0;
  double xi = 0, yi = 0, zi = 0;
  int lat_i = 0, lon_i = 0;
  if (p >= 0) {
    lon_i = p / a.nlat;
    lat_i = p - lon_i * a.nlat;
    xi = a.x[base + p];
    yi = a.x[Np + base + p];
    zi = a.x[2 * Np + base + p];
  }
  const bool compact = a.cell_compact[cell] != 0;
  const double Bc = a.Bcell[cell];
  const double r_eps2 = a.prm.r_eps * a.prm.r_eps;
  const double rc2 = a.prm.rc2_thr, tab_scale = a.prm.tab_scale;
  double ax = 0, ay = 0, az = 0;
  const int chunk = a.chunk_cols * a.nlat;

  for (int col0 = 0; col0 < a.nlon; col0 += a.chunk_cols) {
    const int ncol = min(a.chunk_cols, a.nlon - col0), nsrc = ncol * a.nlat;
    __syncthreads();
    for (int jj = threadIdx.x; jj < nsrc; jj += blockDim.x) {
      const size_t j = base + (size_t)col0 * a.nlat + jj;
      double *r = s_rec + (size_t)jj * PS_REC;
      r[0] = a.x[j];
      r[1] = a.x[Np + j];
      r[2] = a.x[2 * Np + j];
      if (SL) {
        r[3] = a.f[j];
        r[4] = a.f[Np + j];
        r[5] = a.f[2 * Np + j];
      } else {
        r[3] = a.g[j] * Bc;   // slist%Bcoef(j) * slist%g(j,:), ModIntOnRbcs.F90:102-105
        r[4] = a.g[Np + j] * Bc;
        r[5] = a.g[2 * Np + j] * Bc;
        r[6] = a.a3[j];
        r[7] = a.a3[Np + j];
        r[8] = a.a3[2 * Np + j];
      }
    }
    __syncthreads();
    (void)chunk;
    for (int c = 0; c < ncol; c++) {
      const int lon_j = col0 + c;
      int dl = abs(lon_i - lon_j);
      dl = min(dl, a.nlon - dl);
      const unsigned long long bits = s_bits[lat_i * nlonh + dl];
      const double2 *rec = reinterpret_cast<const double2 *>(s_rec + (size_t)c * a.nlat * PS_REC);
#pragma unroll 2
      for (int lat_j = 0; lat_j < a.nlat; lat_j++, rec += PS_REC / 2) {
        const double2 q0 = rec[0], q1 = rec[1];  // (x, y), (z, d0)
        double xx = __dsub_rn(q0.x, xi), yy = __dsub_rn(q0.y, yi), zz = __dsub_rn(q1.x, zi);
        if (!compact) {
          xx = min_image(xx, a.prm.iLb[0], a.prm.Lb[0]);
          yy = min_image(yy, a.prm.iLb[1], a.prm.Lb[1]);
          zz = min_image(zz, a.prm.iLb[2], a.prm.Lb[2]);
        }
        const double r2 = norm2_exact(xx, yy, zz);
        const bool in = valid && !(r2 > rc2) && r2 >= r_eps2;
        if (!__any_sync(FULL_MASK, in)) continue;
        const double2 q2 = rec[2];  // (d1, d2)
        if (in) {
          const double rinv = rsqrt_pos(r2);
          const double s = r2 * rinv * tab_scale;
          const int i = (int)s;
          if (i < RBC3D_NTAB) {
            double om = 1.0;  // 1 - mask
            if ((bits >> lat_j) & 1ull) om = __ldg(a.omm + ((size_t)(lat_i * a.nlat + lat_j)) * nlonh + dl);
            const double fr = s - (double)i;
            const double ir2 = rinv * rinv;
            if (SL) {
              const double t10 = s_tab[2 * i], t20 = s_tab[2 * i + 1], t11 = s_tab[2 * i + 2], t21 = s_tab[2 * i + 3];
              const double e1 = fma(fr, t11 - t10, t10), e2 = fma(fr, t21 - t20, t20);
              const double EA = e1 * rinv * ir2 + e2 * ir2;  // ModEwaldFunc.F90:126-127
              const double EB = e1 * rinv - e2;
              const double fx = q1.y, fy = q2.x, fz = q2.y;
              const double xf = EA * (xx * fx + yy * fy + zz * fz);
              ax += om * (xf * xx + EB * fx);
              ay += om * (xf * yy + EB * fy);
              az += om * (xf * zz + EB * fz);
            } else {
              const double2 q3 = rec[3], q4 = rec[4];  // (n0, n1), (n2, -)
              const double t0 = s_tab[i], t1 = s_tab[i + 1];
              const double e = fma(fr, t1 - t0, t0);
              const double EA = e * ir2 * ir2 * rinv;  // c1 / r^5, ModEwaldFunc.F90:174
              const double qd = om * EA * (xx * q1.y + yy * q2.x + zz * q2.y) * (xx * q3.x + yy * q3.y + zz * q4.x);
              ax += qd * xx;
              ay += qd * yy;
              az += qd * zz;
            }
          }
        }
      }
    }
  }
  if (valid) {
    const double cc = SL ? a.c1 : a.c2;
    const size_t ti = base + p;
    a.acc[ti] += cc * ax;
    a.acc[Np + ti] += cc * ay;
    a.acc[2 * Np + ti] += cc * az;
  }
}

// ---------------------------------------------------------------------------------------------------------
// v3: symmetric evaluation.  The pairs (i <- j) and (j <- i) share the distance, the range test, the table lookup and
// the mask, so the cell's targets are grouped into compact patches of 32 (the same 4 lat x 8 lon patches) and the
// unordered patch pairs (I <= J) are the work units: lane l owns target i_l of patch I in registers and, at step k,
// meets source j_{(l+k) mod 32} of patch J, whose record it reads from a per-warp shared-memory slot (80-byte
// records, conflict-free LDS.128 under rotation).  The contribution to i_l stays in lane l; the contribution to
// j_{l+k} travels in an accumulator that is rotated one lane per step, so after the last step every lane holds the
// sum for its own j.  Patch pairs farther apart than rc + r_I + r_J (bounding spheres) are skipped, the mask table is
// consulted only for patch pairs whose reference-sphere patches can overlap a mask support (cell-independent
// table), and both sides are flushed with FP64 reductions.  Warps draw patches I from a shared counter, largest first.
constexpr int P3_WARPS = 16;

struct Self3Args {
  SelfArgs s;
  const unsigned char *needmask;  // [G][G], cell independent
  const int *cell_list;           // CTA -> cell (null: identity)
  // Few cells: `split` CTAs share a cell and the J loop of a patch I is dealt out in `nj` interleaved parts, so the
  // work units are (I, J = I + jc, I + jc + nj, ...) and CTA `part` of a cell draws units part, part + split, ...
  // (every contribution leaves through an FP64 reduction, so any partition of the patch pairs is valid).  With
  // split = nj = 1 this is one CTA per cell walking I = 0, 1, ... as before.
  int split = 1, nj = 1;
};

template <bool SL>
__global__ void __launch_bounds__(P3_WARPS * 32, 1) k_pair_self3(Self3Args aa) {
  const SelfArgs &a = aa.s;
  extern __shared__ double smem[];
  const int G = a.nwarps_cell;
  const int ntab = SL ? 2 * (RBC3D_NTAB + 1) : (RBC3D_NTAB + 1) + 1;
  double *s_tab = smem;
  double *s_bs = smem + ntab;                        // [G][4] centre, radius
  double *s_rec = s_bs + 4 * ((G + 1) & ~1);         // [P3_WARPS][32][PS_REC]
  __shared__ int s_next;
  const int cslot = blockIdx.x / aa.split, part = blockIdx.x - cslot * aa.split;
  const int cell = aa.cell_list ? aa.cell_list[cslot] : cslot;
  if (!a.cell_active[cell]) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t base = (size_t)cell * a.npc, Np = a.Np;
  const int nlonh = a.nlon / 2 + 1;
  for (int i = threadIdx.x; i <= RBC3D_NTAB; i += blockDim.x) {
    if (SL) {
      s_tab[2 * i] = a.tab_sl[2 * i];
      s_tab[2 * i + 1] = a.tab_sl[2 * i + 1];
    } else {
      s_tab[i] = a.tab_dl[i];
    }
  }
  const bool compact = a.cell_compact[cell] != 0;
  // bounding spheres of the patches (actual, deformed geometry)
  for (int g = warp; g < G; g += P3_WARPS) {
    const int p = a.warp_tgt[g * 32 + lane];
    double x = 0, y = 0, z = 0;
    if (p >= 0) {
      x = a.x[base + p];
      y = a.x[Np + base + p];
      z = a.x[2 * Np + base + p];
    }
    const unsigned m = __ballot_sync(FULL_MASK, p >= 0);
    const double inv = 1.0 / (double)max(__popc(m), 1);
    const double cx = warp_sum(x) * inv, cy = warp_sum(y) * inv, cz = warp_sum(z) * inv;
    double d2 = p >= 0 ? (x - cx) * (x - cx) + (y - cy) * (y - cy) + (z - cz) * (z - cz) : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d2 = fmax(d2, __shfl_xor_sync(FULL_MASK, d2, o));
    if (lane == 0) {
      s_bs[4 * g] = cx;
      s_bs[4 * g + 1] = cy;
      s_bs[4 * g + 2] = cz;
      s_bs[4 * g + 3] = compact ? sqrt(d2) * (1.0 + 1e-12) + 1e-12 : 1e300;  // no culling across periodic wraps
    }
  }
  if (threadIdx.x == 0) s_next = 0;
  __syncthreads();
  const double r_eps2 = a.prm.r_eps * a.prm.r_eps;
  const double rc = a.prm.rc, rc2 = a.prm.rc2_thr, tab_scale = a.prm.tab_scale;
  const double *src_d = SL ? a.f : a.g;
  const double Bc = SL ? 1.0 : a.Bcell[cell];
  double *rec = s_rec + (size_t)warp * 32 * PS_REC;

  const int nunits = G * aa.nj;
  for (;;) {
    int unit = 0;
    if (lane == 0) unit = part + aa.split * atomicAdd(&s_next, 1);
    unit = __shfl_sync(FULL_MASK, unit, 0);
    if (unit >= nunits) break;
    const int I = unit / aa.nj, jc = unit - I * aa.nj;
    const int p_i = a.warp_tgt[I * 32 + lane];
    const bool valid_i = p_i >= 0;
    double xi = 0, yi = 0, zi = 0, d0 = 0, d1 = 0, d2i = 0, n0 = 0, n1 = 0, n2 = 0;
    int lat_i = 0, lon_i = 0;
    if (valid_i) {
      const size_t q = base + p_i;
      xi = a.x[q];
      yi = a.x[Np + q];
      zi = a.x[2 * Np + q];
      d0 = src_d[q] * Bc;
      d1 = src_d[Np + q] * Bc;
      d2i = src_d[2 * Np + q] * Bc;
      if (!SL) {
        n0 = a.a3[q];
        n1 = a.a3[Np + q];
        n2 = a.a3[2 * Np + q];
      }
      lon_i = p_i / a.nlat;
      lat_i = p_i - lon_i * a.nlat;
    }
    const double cIx = s_bs[4 * I], cIy = s_bs[4 * I + 1], cIz = s_bs[4 * I + 2], rI = s_bs[4 * I + 3];
    double ax = 0, ay = 0, az = 0;
    for (int J = I + jc; J < G; J += aa.nj) {
      {
        const double ex = s_bs[4 * J] - cIx, ey = s_bs[4 * J + 1] - cIy, ez = s_bs[4 * J + 2] - cIz;
        const double reach = rc + rI + s_bs[4 * J + 3];
        if (ex * ex + ey * ey + ez * ez > reach * reach) continue;  // warp-uniform
      }
      // stage patch J: lane -> its own record
      __syncwarp();
      {
        const int p_j = a.warp_tgt[J * 32 + lane];
        double *r = rec + lane * PS_REC;
        int meta = -1;
        if (p_j >= 0) {
          const size_t q = base + p_j;
          r[0] = a.x[q];
          r[1] = a.x[Np + q];
          r[2] = a.x[2 * Np + q];
          r[3] = src_d[q] * Bc;
          r[4] = src_d[Np + q] * Bc;
          r[5] = src_d[2 * Np + q] * Bc;
          if (!SL) {
            r[6] = a.a3[q];
            r[7] = a.a3[Np + q];
            r[8] = a.a3[2 * Np + q];
          }
          const int lon_j = p_j / a.nlat;
          meta = (p_j - lon_j * a.nlat) | (lon_j << 8);
        }
        r[9] = __longlong_as_double((long long)meta);
      }
      __syncwarp();
      const bool diag = (I == J);
      const bool nm = aa.needmask[I * G + J] != 0;
      const int kbeg = diag ? 1 : 0, kend = diag ? 17 : 32;
      double bx = 0, by = 0, bz = 0;  // travelling accumulator for j_{lane + k}
      for (int k = kbeg; k < kend; k++) {
        const int jl = (lane + k) & 31;
        const double2 *rj = reinterpret_cast<const double2 *>(rec + jl * PS_REC);
        const double2 q0 = rj[0], q1 = rj[1], q4 = rj[4];  // (x, y), (z, d0), (n2 | -, meta)
        const int meta = (int)__double_as_longlong(q4.y);
        double xx = __dsub_rn(q0.x, xi), yy = __dsub_rn(q0.y, yi), zz = __dsub_rn(q1.x, zi);
        if (!compact) {
          xx = min_image(xx, a.prm.iLb[0], a.prm.Lb[0]);
          yy = min_image(yy, a.prm.iLb[1], a.prm.Lb[1]);
          zz = min_image(zz, a.prm.iLb[2], a.prm.Lb[2]);
        }
        const double r2 = norm2_exact(xx, yy, zz);
        bool in = valid_i && meta >= 0 && !(r2 > rc2) && r2 >= r_eps2;
        if (diag && k == 16 && lane >= 16) in = false;  // {l, l+16} is met from both ends
        if (__any_sync(FULL_MASK, in)) {
          const double2 q2 = rj[2];  // (d1, d2)
          double cxi = 0, cyi = 0, czi = 0, cxj = 0, cyj = 0, czj = 0;
          if (in) {
            const double rinv = rsqrt_pos(r2);
            const double s = r2 * rinv * tab_scale;
            const int it = (int)s;
            if (it < RBC3D_NTAB) {
              double om = 1.0;
              if (nm) {
                const int lat_j = meta & 0xff, lon_j = meta >> 8;
                int dl = abs(lon_i - lon_j);
                dl = min(dl, a.nlon - dl);
                om = __ldg(a.omm + ((size_t)(lat_i * a.nlat + lat_j)) * nlonh + dl);
              }
              const double fr = s - (double)it;
              const double ir2 = rinv * rinv;
              if (SL) {
                const double t10 = s_tab[2 * it], t20 = s_tab[2 * it + 1], t11 = s_tab[2 * it + 2],
                             t21 = s_tab[2 * it + 3];
                const double e1 = fma(fr, t11 - t10, t10), e2 = fma(fr, t21 - t20, t20);
                const double EA = om * (e1 * rinv * ir2 + e2 * ir2), EB = om * (e1 * rinv - e2);
                const double xfj = EA * (xx * q1.y + yy * q2.x + zz * q2.y);
                const double xfi = EA * (xx * d0 + yy * d1 + zz * d2i);
                cxi = xfj * xx + EB * q1.y;
                cyi = xfj * yy + EB * q2.x;
                czi = xfj * zz + EB * q2.y;
                cxj = xfi * xx + EB * d0;
                cyj = xfi * yy + EB * d1;
                czj = xfi * zz + EB * d2i;
              } else {
                const double2 q3 = rj[3];  // (n0, n1)
                const double t0 = s_tab[it], t1 = s_tab[it + 1];
                const double e = fma(fr, t1 - t0, t0);
                const double EA = om * e * ir2 * ir2 * rinv;
                const double qj = EA * (xx * q1.y + yy * q2.x + zz * q2.y) * (xx * q3.x + yy * q3.y + zz * q4.x);
                const double qi = -EA * (xx * d0 + yy * d1 + zz * d2i) * (xx * n0 + yy * n1 + zz * n2);
                cxi = qj * xx;
                cyi = qj * yy;
                czi = qj * zz;
                cxj = qi * xx;
                cyj = qi * yy;
                czj = qi * zz;
              }
            }
          }
          ax += cxi;
          ay += cyi;
          az += czi;
          bx += cxj;
          by += cyj;
          bz += czj;
        }
        if (k + 1 < kend) {  // hand the travelling accumulator to the lane that meets the same j next
          const int from = (lane + 1) & 31;
          bx = __shfl_sync(FULL_MASK, bx, from);
          by = __shfl_sync(FULL_MASK, by, from);
          bz = __shfl_sync(FULL_MASK, bz, from);
        }
      }
      // after the last step lane l holds the sum for j_{(l + kend - 1) mod 32}: deliver and flush
      {
        const int from = (lane - (kend - 1)) & 31;
        bx = __shfl_sync(FULL_MASK, bx, from);
        by = __shfl_sync(FULL_MASK, by, from);
        bz = __shfl_sync(FULL_MASK, bz, from);
        const int p_j = a.warp_tgt[J * 32 + lane];
        if (p_j >= 0 && a.active[base + p_j]) {
          const double cc = SL ? a.c1 : a.c2;
          const size_t tj = base + p_j;
          atomicAdd(a.acc + tj, cc * bx);
          atomicAdd(a.acc + Np + tj, cc * by);
          atomicAdd(a.acc + 2 * Np + tj, cc * bz);
        }
      }
    }
    if (valid_i && a.active[base + p_i]) {
      const double cc = SL ? a.c1 : a.c2;
      const size_t ti = base + p_i;
      atomicAdd(a.acc + ti, cc * ax);
      atomicAdd(a.acc + Np + ti, cc * ay);
      atomicAdd(a.acc + 2 * Np + ti, cc * az);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------
// v4: the symmetric kernel with everything that does not depend on the density hoisted out of the GMRES matvec.
// For an unordered pair the distance, the range test, the table lookup, the rsqrt and the mask give ONE number,
// coef = (1 - mask) EA(|xx|), shared by both directions; it only changes with the geometry, i.e. once per time step.
// k_pc_build walks the patch pairs exactly like k_pair_self3 and records, per patch pair, the 32-bit set of rotation
// steps in which at least one lane has an in-range pair (scan pass), then the coefficients of those steps, 32
// doubles = one coalesced 256-byte row per step (fill pass).  The matvec (k_pair_self_cached) streams the rows once
// and is left with 3 subtractions, 4 dot products and 10 multiply-adds per unordered pair -- no table, no rsqrt, no
// mask, no range test, no branch.  8 B per lane and step: ~16 MB per 36 x 72 cell at rc = 1.2; cells that do not fit
// in device memory keep the direct kernel.
struct PcArgs {
  SelfArgs s;
  const unsigned char *needmask;
  const int *cell_list;      // slot -> cell
  unsigned *mask;            // [slot][G][G]
  int *cnt;                  // scan: steps of (slot, I); fill / apply: first step of (slot, I)
  double *coef;              // [step][32]
};

template <bool FILL>
__global__ void __launch_bounds__(P3_WARPS * 32, 1) k_pc_build(PcArgs aa) {
  const SelfArgs &a = aa.s;
  extern __shared__ double smem[];
  const int G = a.nwarps_cell;
  double *s_tab = smem;                                  // DL table (fill pass only)
  double *s_bs = smem + (FILL ? RBC3D_NTAB + 2 : 0);     // [G][4] centre, radius
  double *s_rec = s_bs + 4 * ((G + 1) & ~1);             // [P3_WARPS][32][4] x, y, z, meta
  __shared__ int s_next;
  const int slot = blockIdx.x, cell = aa.cell_list[slot];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t base = (size_t)cell * a.npc, Np = a.Np;
  const int nlonh = a.nlon / 2 + 1;
  if (FILL)
    for (int i = threadIdx.x; i <= RBC3D_NTAB; i += blockDim.x) s_tab[i] = a.tab_dl[i];
  const bool compact = a.cell_compact[cell] != 0;
  for (int g = warp; g < G; g += P3_WARPS) {  // bounding spheres, as in k_pair_self3
    const int p = a.warp_tgt[g * 32 + lane];
    double x = 0, y = 0, z = 0;
    if (p >= 0) {
      x = a.x[base + p];
      y = a.x[Np + base + p];
      z = a.x[2 * Np + base + p];
    }
    const unsigned m = __ballot_sync(FULL_MASK, p >= 0);
    const double inv = 1.0 / (double)max(__popc(m), 1);
    const double cx = warp_sum(x) * inv, cy = warp_sum(y) * inv, cz = warp_sum(z) * inv;
    double d2 = p >= 0 ? (x - cx) * (x - cx) + (y - cy) * (y - cy) + (z - cz) * (z - cz) : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d2 = fmax(d2, __shfl_xor_sync(FULL_MASK, d2, o));
    if (lane == 0) {
      s_bs[4 * g] = cx;
      s_bs[4 * g + 1] = cy;
      s_bs[4 * g + 2] = cz;
      s_bs[4 * g + 3] = compact ? sqrt(d2) * (1.0 + 1e-12) + 1e-12 : 1e300;
    }
  }
  if (threadIdx.x == 0) s_next = 0;
  __syncthreads();
  const double r_eps2 = a.prm.r_eps * a.prm.r_eps;
  const double rc = a.prm.rc, rc2 = a.prm.rc2_thr, tab_scale = a.prm.tab_scale;
  double *rec = s_rec + (size_t)warp * 32 * 4;
  for (;;) {
    int I = 0;
    if (lane == 0) I = atomicAdd(&s_next, 1);
    I = __shfl_sync(FULL_MASK, I, 0);
    if (I >= G) break;
    const int p_i = a.warp_tgt[I * 32 + lane];
    const bool valid_i = p_i >= 0;
    double xi = 0, yi = 0, zi = 0;
    int lat_i = 0, lon_i = 0;
    if (valid_i) {
      const size_t q = base + p_i;
      xi = a.x[q];
      yi = a.x[Np + q];
      zi = a.x[2 * Np + q];
      lon_i = p_i / a.nlat;
      lat_i = p_i - lon_i * a.nlat;
    }
    const double cIx = s_bs[4 * I], cIy = s_bs[4 * I + 1], cIz = s_bs[4 * I + 2], rI = s_bs[4 * I + 3];
    unsigned *mrow = aa.mask + ((size_t)slot * G + I) * G;
    double *out = FILL ? aa.coef + (size_t)aa.cnt[slot * G + I] * 32 + lane : nullptr;
    int nsteps = 0;
    for (int J = I; J < G; J++) {
      unsigned msk = 0;
      if (FILL) {
        msk = mrow[J];
        if (!msk) continue;
      } else {
        const double ex = s_bs[4 * J] - cIx, ey = s_bs[4 * J + 1] - cIy, ez = s_bs[4 * J + 2] - cIz;
        const double reach = rc + rI + s_bs[4 * J + 3];
        if (ex * ex + ey * ey + ez * ez > reach * reach) {  // warp-uniform
          if (lane == 0) mrow[J] = 0;
          continue;
        }
      }
      __syncwarp();
      {
        const int p_j = a.warp_tgt[J * 32 + lane];
        double *r = rec + lane * 4;
        int meta = -1;
        if (p_j >= 0) {
          const size_t q = base + p_j;
          r[0] = a.x[q];
          r[1] = a.x[Np + q];
          r[2] = a.x[2 * Np + q];
          const int lon_j = p_j / a.nlat;
          meta = (p_j - lon_j * a.nlat) | (lon_j << 8);
        }
        r[3] = __longlong_as_double((long long)meta);
      }
      __syncwarp();
      const bool diag = (I == J);
      const bool nm = aa.needmask[I * G + J] != 0;
      const int kbeg = diag ? 1 : 0, kend = diag ? 17 : 32;
      for (int k = kbeg; k < kend; k++) {
        if (FILL && !((msk >> k) & 1)) continue;
        const int jl = (lane + k) & 31;
        const double2 *rj = reinterpret_cast<const double2 *>(rec + jl * 4);
        const double2 q0 = rj[0], q1 = rj[1];
        const int meta = (int)__double_as_longlong(q1.y);
        double xx = __dsub_rn(q0.x, xi), yy = __dsub_rn(q0.y, yi), zz = __dsub_rn(q1.x, zi);
        if (!compact) {
          xx = min_image(xx, a.prm.iLb[0], a.prm.Lb[0]);
          yy = min_image(yy, a.prm.iLb[1], a.prm.Lb[1]);
          zz = min_image(zz, a.prm.iLb[2], a.prm.Lb[2]);
        }
        const double r2 = norm2_exact(xx, yy, zz);
        bool in = valid_i && meta >= 0 && !(r2 > rc2) && r2 >= r_eps2;
        if (diag && k == 16 && lane >= 16) in = false;  // {l, l+16} is met from both ends
        if (!FILL) {
          if (__any_sync(FULL_MASK, in)) {
            msk |= 1u << k;
            nsteps++;
          }
        } else {
          double coef = 0.0;
          if (in) {
            const double rinv = rsqrt_pos(r2);
            const double sc = r2 * rinv * tab_scale;
            const int it = (int)sc;
            if (it < RBC3D_NTAB) {
              double om = 1.0;
              if (nm) {
                const int lat_j = meta & 0xff, lon_j = meta >> 8;
                int dl = abs(lon_i - lon_j);
                dl = min(dl, a.nlon - dl);
                om = __ldg(a.omm + ((size_t)(lat_i * a.nlat + lat_j)) * nlonh + dl);
              }
              const double fr = sc - (double)it;
              const double ir2 = rinv * rinv;
              const double t0 = s_tab[it], t1 = s_tab[it + 1];
              const double e = fma(fr, t1 - t0, t0);
              coef = om * e * ir2 * ir2 * rinv;
            }
          }
          *out = coef;
          out += 32;
        }
      }
      if (!FILL && lane == 0) mrow[J] = msk;
    }
    if (!FILL && lane == 0) aa.cnt[slot * G + I] = nsteps;
  }
}

// streaming 8-byte load of a coefficient row element: read once per matvec, no L1 allocation
__device__ __forceinline__ double ld_stream1(const double *p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

constexpr int PC_WARPS = 16;

__global__ void __launch_bounds__(PC_WARPS * 32, 2) k_pair_self_cached(PcArgs aa) {
  const SelfArgs &a = aa.s;
  __shared__ __align__(16) double s_rec[PC_WARPS][32 * PS_REC];
  __shared__ int s_next;
  const int G = a.nwarps_cell;
  const int slot = blockIdx.x, cell = aa.cell_list[slot];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t base = (size_t)cell * a.npc, Np = a.Np;
  const bool compact = a.cell_compact[cell] != 0;
  const double Bc = a.Bcell[cell];
  if (threadIdx.x == 0) s_next = 0;
  __syncthreads();
  double *rec = s_rec[warp];
  for (;;) {
    int I = 0;
    if (lane == 0) I = atomicAdd(&s_next, 1);
    I = __shfl_sync(FULL_MASK, I, 0);
    if (I >= G) break;
    const unsigned *mrow = aa.mask + ((size_t)slot * G + I) * G;
    const double *cp = aa.coef + (size_t)aa.cnt[slot * G + I] * 32 + lane;
    const int p_i = a.warp_tgt[I * 32 + lane];
    const bool valid_i = p_i >= 0;
    double xi = 0, yi = 0, zi = 0, d0 = 0, d1 = 0, d2i = 0, n0 = 0, n1 = 0, n2 = 0;
    if (valid_i) {
      const size_t q = base + p_i;
      xi = a.x[q];
      yi = a.x[Np + q];
      zi = a.x[2 * Np + q];
      d0 = a.g[q] * Bc;
      d1 = a.g[Np + q] * Bc;
      d2i = a.g[2 * Np + q] * Bc;
      n0 = a.a3[q];
      n1 = a.a3[Np + q];
      n2 = a.a3[2 * Np + q];
    }
    double ax = 0, ay = 0, az = 0;
    // masks of the row, 32 patch pairs per load
    for (int J0 = I & ~31; J0 < G; J0 += 32) {
      const unsigned mymask = (J0 + lane >= I && J0 + lane < G) ? mrow[J0 + lane] : 0u;
      unsigned todo = __ballot_sync(FULL_MASK, mymask != 0);
      while (todo) {
        const int jj = __ffs(todo) - 1;
        todo &= todo - 1;
        const int J = J0 + jj;
        unsigned msk = __shfl_sync(FULL_MASK, mymask, jj);
        double cnext = ld_stream1(cp);  // first step's coefficients in flight while patch J is staged
        __syncwarp();
        const int p_j = a.warp_tgt[J * 32 + lane];
        {
          double *r = rec + lane * PS_REC;
          if (p_j >= 0) {
            const size_t q = base + p_j;
            r[0] = a.x[q];
            r[1] = a.x[Np + q];
            r[2] = a.x[2 * Np + q];
            r[3] = a.g[q] * Bc;
            r[4] = a.g[Np + q] * Bc;
            r[5] = a.g[2 * Np + q] * Bc;
            r[6] = a.a3[q];
            r[7] = a.a3[Np + q];
            r[8] = a.a3[2 * Np + q];
          } else {  // coefficients of missing points are zero; keep the arithmetic finite
#pragma unroll
            for (int u = 0; u < 9; u++) r[u] = 0.0;
          }
        }
        __syncwarp();
        double bx = 0, by = 0, bz = 0;  // travelling accumulator, as in k_pair_self3
        int kprev = -1;
        while (msk) {
          const int k = __ffs(msk) - 1;
          msk &= msk - 1;
          const double coef = cnext;
          cp += 32;
          if (msk) cnext = ld_stream1(cp);
          if (kprev >= 0) {  // the lane that meets the same j at step k sits k - kprev lanes below
            const int from = (lane + (k - kprev)) & 31;
            bx = __shfl_sync(FULL_MASK, bx, from);
            by = __shfl_sync(FULL_MASK, by, from);
            bz = __shfl_sync(FULL_MASK, bz, from);
          }
          kprev = k;
          const int jl = (lane + k) & 31;
          const double2 *rj = reinterpret_cast<const double2 *>(rec + jl * PS_REC);
          const double2 q0 = rj[0], q1 = rj[1], q2 = rj[2], q3 = rj[3], q4 = rj[4];
          double xx = q0.x - xi, yy = q0.y - yi, zz = q1.x - zi;
          if (!compact) {
            xx = min_image(xx, a.prm.iLb[0], a.prm.Lb[0]);
            yy = min_image(yy, a.prm.iLb[1], a.prm.Lb[1]);
            zz = min_image(zz, a.prm.iLb[2], a.prm.Lb[2]);
          }
          const double qj = coef * (xx * q1.y + yy * q2.x + zz * q2.y) * (xx * q3.x + yy * q3.y + zz * q4.x);
          const double qi = -coef * (xx * d0 + yy * d1 + zz * d2i) * (xx * n0 + yy * n1 + zz * n2);
          ax = fma(qj, xx, ax);
          ay = fma(qj, yy, ay);
          az = fma(qj, zz, az);
          bx = fma(qi, xx, bx);
          by = fma(qi, yy, by);
          bz = fma(qi, zz, bz);
        }
        {  // lane l holds the sum for j_{(l + kprev) mod 32}: deliver and flush
          const int from = (lane - kprev) & 31;
          bx = __shfl_sync(FULL_MASK, bx, from);
          by = __shfl_sync(FULL_MASK, by, from);
          bz = __shfl_sync(FULL_MASK, bz, from);
          if (p_j >= 0 && a.active[base + p_j]) {
            const size_t tj = base + p_j;
            atomicAdd(a.acc + tj, a.c2 * bx);
            atomicAdd(a.acc + Np + tj, a.c2 * by);
            atomicAdd(a.acc + 2 * Np + tj, a.c2 * bz);
          }
        }
      }
    }
    if (valid_i && a.active[base + p_i]) {
      const size_t ti = base + p_i;
      atomicAdd(a.acc + ti, a.c2 * ax);
      atomicAdd(a.acc + Np + ti, a.c2 * ay);
      atomicAdd(a.acc + 2 * Np + ti, a.c2 * az);
    }
  }
}

// per cell: is the extent of the cell below half a box in every direction (then nint((xj-xi)/L) = 0 exactly)?
__global__ void __launch_bounds__(256) k_cell_compact(int npc, int Np, const double *__restrict__ x, Params prm,
                                                      unsigned char *__restrict__ compact) {
  const int cell = blockIdx.x;
  __shared__ double s_mn[3][8], s_mx[3][8];
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int i = threadIdx.x; i < npc; i += blockDim.x)
#pragma unroll
    for (int d = 0; d < 3; d++) {
      const double v = x[(size_t)d * Np + (size_t)cell * npc + i];
      mn[d] = fmin(mn[d], v);
      mx[d] = fmax(mx[d], v);
    }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = fmin(mn[d], __shfl_xor_sync(FULL_MASK, mn[d], o));
      mx[d] = fmax(mx[d], __shfl_xor_sync(FULL_MASK, mx[d], o));
    }
    if (lane == 0) {
      s_mn[d][warp] = mn[d];
      s_mx[d][warp] = mx[d];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    bool ok = true;
    for (int d = 0; d < 3; d++) {
      double a = s_mn[d][0], b = s_mx[d][0];
      for (int w = 1; w < 8; w++) {
        a = fmin(a, s_mn[d][w]);
        b = fmax(b, s_mx[d][w]);
      }
      ok = ok && ((b - a) * prm.iLb[d] < 0.49);
    }
    compact[cell] = ok ? 1 : 0;
  }
}

// cell-independent tables (rbc3d_cells_set_mesh): target -> (warp, lane) map and the mask bit sets
int pairself_mesh_prepare(rbc3d_ctx *c, const std::vector<double> &omm) {
  Cells &C = c->cells;
  C.ps_ok = false;
  const int nlat = C.nlat, nlon = C.nlon, nlonh = nlon / 2 + 1;
  if (nlat > 64) return RBC3D_OK;  // bit sets are 64 wide; the cell-list kernel handles such meshes
  // warps = compact blocks of 4 lat x 8 lon targets
  const int nbl = (nlat + 3) / 4, nbn = (nlon + 7) / 8;
  std::vector<int> wt((size_t)nbl * nbn * 32, -1);
  for (int bl = 0; bl < nbl; bl++)
    for (int bn = 0; bn < nbn; bn++)
      for (int lane = 0; lane < 32; lane++) {
        const int ilat = bl * 4 + (lane & 3), ilon = bn * 8 + (lane >> 2);
        if (ilat < nlat && ilon < nlon) wt[((size_t)bl * nbn + bn) * 32 + lane] = ilon * nlat + ilat;
      }
  std::vector<unsigned long long> bits((size_t)nlat * nlonh, 0ull);
  for (int i = 0; i < nlat; i++)
    for (int j = 0; j < nlat; j++)
      for (int dl = 0; dl < nlonh; dl++)
        if (omm[((size_t)i * nlat + j) * nlonh + dl] != 1.0) bits[(size_t)i * nlonh + dl] |= 1ull << j;
  C.ps_nwarps = nbl * nbn;
  // patch pairs whose mask supports can overlap
  const int G = nbl * nbn;
  std::vector<unsigned char> nm((size_t)G * G, 0);
  for (int I = 0; I < G; I++)
    for (int J = 0; J < G; J++) {
      bool any = false;
      for (int li = 0; li < 32 && !any; li++) {
        const int pi = wt[(size_t)I * 32 + li];
        if (pi < 0) continue;
        for (int lj = 0; lj < 32 && !any; lj++) {
          const int pj = wt[(size_t)J * 32 + lj];
          if (pj < 0) continue;
          int dl = abs(pi / nlat - pj / nlat);
          dl = std::min(dl, nlon - dl);
          any = omm[((size_t)(pi % nlat) * nlat + (pj % nlat)) * nlonh + dl] != 1.0;
        }
      }
      nm[(size_t)I * G + J] = any ? 1 : 0;
    }
  RBC_TRY(C.ps_needmask.resize(nm.size()));
  CUDA_TRY(cudaMemcpyAsync(C.ps_needmask.p, nm.data(), nm.size(), cudaMemcpyHostToDevice, c->stream));
  RBC_TRY(C.ps_warp_tgt.resize(wt.size()));
  RBC_TRY(C.ps_maskbits.resize(bits.size()));
  CUDA_TRY(cudaMemcpyAsync(C.ps_warp_tgt.p, wt.data(), sizeof(int) * wt.size(), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(C.ps_maskbits.p, bits.data(), sizeof(unsigned long long) * bits.size(),
                           cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  C.ps_ok = true;
  return RBC3D_OK;
}

// geometry time
int pairself_geometry_prepare(rbc3d_ctx *c) {
  Cells &C = c->cells;
  if (C.ncell == 0) return RBC3D_OK;
  RBC_TRY(C.ps_compact.resize(C.ncell));
  k_cell_compact<<<C.ncell, 256, 0, c->stream>>>(C.npc, C.Np, C.x.p, c->prm, C.ps_compact.p);
  KERNEL_CHECK();
  c->launches++;
  return RBC3D_OK;
}

static void fill_self_args(rbc3d_ctx *c, SelfArgs &a, const int *active, double *acc, double c1, double c2) {
  Cells &C = c->cells;
  a.prm = c->prm;
  a.ncell = C.ncell;
  a.npc = C.npc;
  a.nlat = C.nlat;
  a.nlon = C.nlon;
  a.Np = C.Np;
  a.x = C.x.p;
  a.a3 = C.a3.p;
  a.f = C.f.p;
  a.g = C.g.p;
  a.Bcell = C.B.p;
  a.warp_tgt = C.ps_warp_tgt.p;
  a.nwarps_cell = C.ps_nwarps;
  a.ctas_per_cell = (C.ps_nwarps + PS_WARPS - 1) / PS_WARPS;
  a.maskbits = C.ps_maskbits.p;
  a.omm = C.omm.p;
  a.tab_sl = c->tab_sl.p;
  a.tab_dl = c->tab_dl.p;
  a.active = active;
  a.cell_active = C.sg_cell_active.p;
  a.cell_compact = C.ps_compact.p;
  a.c1 = c1;
  a.c2 = c2;
  a.acc = acc;
  a.chunk_cols = 0;
}

// geometry time, after singular_prepare (the singular cache has the first call on device memory): coefficient cache
// of the symmetric double-layer pair sum for as many active cells as fit
int pairself_cache_prepare(rbc3d_ctx *c) {
  Cells &C = c->cells;
  C.pc_ok = false;
  C.pc_ncached = 0;
  if (!C.ps_ok || C.Np == 0 || c->pair_self_mode != 3 || C.sg_nactive == 0) return RBC3D_OK;
  const int G = C.ps_nwarps, nslot = C.sg_nactive;
  if (C.pc_mask.resize((size_t)nslot * G * G) != RBC3D_OK) return RBC3D_OK;
  if (C.pc_cnt.resize((size_t)nslot * G + 1) != RBC3D_OK) return RBC3D_OK;
  PcArgs aa;
  fill_self_args(c, aa.s, nullptr, nullptr, 0.0, 0.0);
  aa.needmask = C.ps_needmask.p;
  aa.cell_list = C.sg_active_list.p;
  aa.mask = C.pc_mask.p;
  aa.cnt = C.pc_cnt.p;
  aa.coef = nullptr;
  const size_t smem_scan = sizeof(double) * (4 * ((G + 1) & ~1) + (size_t)P3_WARPS * 32 * 4);
  const size_t smem_fill = smem_scan + sizeof(double) * (RBC3D_NTAB + 2);
  CUDA_TRY(cudaFuncSetAttribute(k_pc_build<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scan));
  CUDA_TRY(cudaFuncSetAttribute(k_pc_build<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fill));
  k_pc_build<false><<<nslot, P3_WARPS * 32, smem_scan, c->stream>>>(aa);
  KERNEL_CHECK();
  c->launches++;
  // steps per slot (host decides how many cells fit), then first step of every (slot, I)
  std::vector<int> cnt((size_t)nslot * G);
  CUDA_TRY(cudaMemcpyAsync(cnt.data(), C.pc_cnt.p, sizeof(int) * cnt.size(), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  size_t free_b = 0, total_b = 0;
  CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
  // the buffer is grow-only: what it already holds counts as available; keep room for the density splines, the PME
  // work arrays and lists that are allocated lazily
  // (built on first use, the density splines usually exist by now: only what is still to come is held back)
  const size_t spl = (size_t)C.ncell * 12 * 2 * C.nlat * C.nlon * 8;
  const size_t reserve = (C.spG.n ? 0 : spl) + (C.spGi.n ? 0 : spl) + ((size_t)6 << 30);
  size_t budget = free_b + C.pc_coef.n * sizeof(double);
  budget = budget > reserve ? budget - reserve : 0;
  long long max_cells = nslot;
  if (const char *e = getenv("RBC3D_PAIR_CACHE_MAX_CELLS")) max_cells = atoll(e);
  long long steps = 0;
  int ncached = 0;
  for (int sl = 0; sl < nslot && sl < max_cells; sl++) {
    long long ssum = 0;
    for (int I = 0; I < G; I++) ssum += cnt[(size_t)sl * G + I];
    if ((size_t)(steps + ssum) * 256 > budget || steps + ssum > 2000000000ll) break;
    steps += ssum;
    ncached++;
  }
  if (ncached == 0) return RBC3D_OK;
  {  // exclusive scan on the host copy (nslot * G ints), uploaded as the offsets of the fill / apply kernels
    long long run = 0;
    for (size_t i = 0; i < (size_t)ncached * G; i++) {
      const int v = cnt[i];
      cnt[i] = (int)run;
      run += v;
    }
    CUDA_TRY(cudaMemcpyAsync(C.pc_cnt.p, cnt.data(), sizeof(int) * (size_t)ncached * G, cudaMemcpyHostToDevice, c->stream));
  }
  if (C.pc_coef.resize((size_t)steps * 32 + 32) != RBC3D_OK) return RBC3D_OK;
  aa.coef = C.pc_coef.p;
  k_pc_build<true><<<ncached, P3_WARPS * 32, smem_fill, c->stream>>>(aa);
  KERNEL_CHECK();
  c->launches++;
  CUDA_TRY(cudaStreamSynchronize(c->stream));  // cnt (host vector) must outlive the upload
  C.pc_ncached = ncached;
  C.pc_rows = steps;
  C.pc_ok = true;
  return RBC3D_OK;
}

bool pairself_available(rbc3d_ctx *c, const TargetList &t) {
  return c->cells.ps_ok && t.kind == RBC3D_TL_CELLS && c->pair_self_mode != 0;
}

int pairself_apply(rbc3d_ctx *c, TargetList &t, double c1, double c2) {
  Cells &C = c->cells;
  SelfArgs a;
  fill_self_args(c, a, t.active.p, t.acc.p, c1, c2);
  const int nlonh = C.nlon / 2 + 1;
  const int nbits = (C.nlat * nlonh + 1) & ~1;
  const int grid = C.ncell * a.ctas_per_cell;
  if (c->pair_self_mode == 1 || c->pair_self_mode == 3) {  // symmetric patch-pair kernels
    Self3Args aa;
    aa.s = a;
    aa.needmask = C.ps_needmask.p;
    aa.cell_list = nullptr;
    int ncells_direct = C.ncell;
    if (c->pair_self_mode == 3 && C.pc_pending && c1 == 0 && c2 != 0) {
      C.pc_pending = false;
      RBC_TRY(pairself_cache_prepare(c));
    }
    const bool cached = c->pair_self_mode == 3 && C.pc_ok && C.pc_ncached > 0 && c1 == 0 && c2 != 0;
    if (cached) {  // double layer alone (the GMRES matvec): stream the geometry cache, direct kernel for the rest
      PcArgs pa;
      pa.s = a;
      pa.needmask = C.ps_needmask.p;
      pa.cell_list = C.sg_active_list.p;
      pa.mask = C.pc_mask.p;
      pa.cnt = C.pc_cnt.p;
      pa.coef = C.pc_coef.p;
      k_pair_self_cached<<<C.pc_ncached, PC_WARPS * 32, 0, c->stream>>>(pa);
      KERNEL_CHECK();
      c->launches++;
      aa.cell_list = C.sg_active_list.p + C.pc_ncached;
      ncells_direct = C.sg_nactive - C.pc_ncached;
      if (ncells_direct <= 0) return RBC3D_OK;
    }
    // fewer cells than SMs: several CTAs per cell, about one work unit per warp (examples/minicase has 2 cells)
    if (ncells_direct < c->sm_count) {
      aa.split = std::min(64, (2 * c->sm_count + ncells_direct - 1) / ncells_direct);
      aa.nj = std::max(1, std::min(8, (P3_WARPS * aa.split + C.ps_nwarps - 1) / C.ps_nwarps));
    }
    for (int pass = 0; pass < 2; pass++) {
      const bool sl = pass == 0;
      if (sl ? (c1 == 0) : (c2 == 0)) continue;
      const int ntab = sl ? 2 * (RBC3D_NTAB + 1) : (RBC3D_NTAB + 1) + 1;
      const size_t smem = sizeof(double) * ((size_t)ntab + 4 * ((C.ps_nwarps + 1) & ~1) + (size_t)P3_WARPS * 32 * PS_REC);
      if (sl) {
        CUDA_TRY(cudaFuncSetAttribute(k_pair_self3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_pair_self3<true><<<ncells_direct * aa.split, P3_WARPS * 32, smem, c->stream>>>(aa);
      } else {
        CUDA_TRY(cudaFuncSetAttribute(k_pair_self3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_pair_self3<false><<<ncells_direct * aa.split, P3_WARPS * 32, smem, c->stream>>>(aa);
      }
      KERNEL_CHECK();
      c->launches++;
    }
    return RBC3D_OK;
  }
  for (int pass = 0; pass < 2; pass++) {
    const bool sl = pass == 0;
    if (sl ? (c1 == 0) : (c2 == 0)) continue;
    const int ntab = sl ? 2 * (RBC3D_NTAB + 1) : (RBC3D_NTAB + 1) + 1;
    const size_t fixed = sizeof(double) * ((size_t)ntab + nbits);
    const size_t budget = 224 * 1024 - fixed;
    int cols = (int)(budget / (sizeof(double) * PS_REC * C.nlat));
    cols = std::min(cols, std::min(C.nlon, PS_CHUNK_MAX / C.nlat > 0 ? PS_CHUNK_MAX / C.nlat : 1));
    if (cols < 1) cols = 1;
    a.chunk_cols = cols;
    const size_t smem = fixed + sizeof(double) * PS_REC * (size_t)cols * C.nlat;
    if (sl) {
      CUDA_TRY(cudaFuncSetAttribute(k_pair_self<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_pair_self<true><<<grid, PS_WARPS * 32, smem, c->stream>>>(a);
    } else {
      CUDA_TRY(cudaFuncSetAttribute(k_pair_self<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_pair_self<false><<<grid, PS_WARPS * 32, smem, c->stream>>>(a);
    }
    KERNEL_CHECK();
    c->launches++;
  }
  return RBC3D_OK;
}

}  // namespace rbc3d
