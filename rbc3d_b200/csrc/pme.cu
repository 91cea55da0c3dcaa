// pme.cu -- smooth particle-mesh Ewald long-range part: PME_Init / PME_Distrib_Source / PME_Transform /
// PME_Add_Interp_Vel (ModPME.F90:58-338, 405-489) with the slab FFT of ModPFFTW.F90 replaced by cuFFT.
//
// Mesh layout in HBM: real meshes [comp][Nz][Ny][Nx] (x fastest), spectra [comp][Nz][Ny][Nx/2+1] (cuFFT D2Z).
// Components: 0..2 single-layer force density, 3..8 the SYMMETRIC part of the double-layer tensor
// (xx,yy,zz,xy,xz,yz): the k-space multiplier (ModPME.F90:193-202) only sees q tr(T) + q^T T + T q = q tr(S) + 2 S q
// and q^T T q = q^T S q, so 6 transforms replace the reference's 9.
//
// FFT conventions (ModPFFTW.F90:110-114): the reference transforms (x,y) with exp(-i..) and z with exp(+i..) and
// stores q = (i/L1, jhat/L2, -khat/L3).  With a standard all-negative D2Z the coefficient the reference keeps at z
// index k sits at (Nz-k) mod Nz, which turns its wave vector into q3 = +k/L3 for k <= Nz/2 and (k-Nz)/L3 above, while
// q2 keeps "index >= Ny/2 is negative".  FFTW's c2r ignores the imaginary part of the x = 0 and x = Nx/2 bins after the
// y/z transforms; the scaling kernel reproduces that by Hermitian-symmetrising those two planes.
#include <cmath>
#include <complex>
#include <cstdlib>

#include "device_math.cuh"
#include "rbc3d_internal.h"

namespace rbc3d {

constexpr int PME_BLK = 4;   // edge of a PME block in mesh cells (y, z; and x for the P != 8 kernels)
constexpr int SPR_BX = 8, SPR_BY = 4, SPR_BZ = 4;  // source blocks of the P = 8 spreading kernel
constexpr int PME_PMAX = 8;  // largest supported B-spline support (PBspln_Ewd = 8 in every example)
constexpr int SPREAD_CHUNK = 32;
constexpr int SW_XR = 8;     // mesh cells per x run of the walking spread kernel (k_spread_walk)

int pme_block_edge() { return PME_BLK; }
static int pme_transform_slab(rbc3d_ctx *c);
static int pme_halo_exchange(rbc3d_ctx *c, TargetList &t);

void t_begin(rbc3d_ctx *c, int s) {
  if (c->quiet) return;
  cudaEventRecord(c->ev[2 * s], c->stream);
  c->ev_used[s] = true;
}
void t_end(rbc3d_ctx *c, int s) {
  if (!c->quiet) cudaEventRecord(c->ev[2 * s + 1], c->stream);
}

// P = 8 has two spreading kernels.  The pencil walk (k_spread_walk) is the fast one when every pencil has work; a
// short list (the 2 cells or the 2404 wall triangles of examples/minicase) leaves most pencils empty and turns the
// few occupied ones into long serial chains, where the source-block kernel (k_spread8, 9 times more work units on
// the same sources) is several times faster.  The sort key of the list follows the choice.
void pme_spread_mode(rbc3d_ctx *c, CellList &L, int n) {
  const Pme &pm = c->pme;
  const Params &p = c->prm;
  const int sb[3] = {SPR_BX, SPR_BY, SPR_BZ};
  L.swalk = (p.P == 8) && (pm.swalk_mode < 0 ? n >= pm.swalk_min : pm.swalk_mode == 0);
  for (int d = 0; d < 3; d++) {
    L.sblk[d] = (p.P == 8) ? (L.swalk ? (d == 0 ? SW_XR : 1) : sb[d]) : PME_BLK;
    L.nsblk[d] = (p.Nb[d] + L.sblk[d] - 1) / L.sblk[d];
  }
}

int pme_init(rbc3d_ctx *c) {
  Pme &pm = c->pme;
  const Params &p = c->prm;
  if (p.P < 2 || p.P > PME_PMAX) {
    set_error("PBspln = %d unsupported (2..%d)", p.P, PME_PMAX);
    return RBC3D_EINVAL;
  }
  pm.Nx = p.Nb[0];
  pm.Ny = p.Nb[1];
  pm.Nz = p.Nb[2];
  pm.Nxh = pm.Nx / 2 + 1;
  pm.G = (size_t)pm.Nx * pm.Ny * pm.Nz;
  pm.M = (size_t)pm.Nxh * pm.Ny * pm.Nz;
  pm.swalk_mode = -1;
  if (const char *e = getenv("RBC3D_SPREAD_BLOCKS")) pm.swalk_mode = atoi(e) == 1 ? 1 : 0;
  if (const char *e = getenv("RBC3D_SPREAD_WALK_MIN")) pm.swalk_min = atoi(e);
  if (const char *e = getenv("RBC3D_INTERP_DIRECT_MAX")) pm.interp_direct_max = atoi(e);
  for (int d = 0; d < 3; d++) {
    pm.iblk[d] = (p.P == 8) ? 1 : PME_BLK;  // P = 8: targets keyed by mesh cell for the column walk (k_interp_walk)
    pm.nblk[d] = (p.Nb[d] + pm.iblk[d] - 1) / pm.iblk[d];
  }
  pm.walk = (p.P == 8);
  RBC_TRY(pm.src.resize(9 * pm.G));
  RBC_TRY(pm.srcC.resize(9 * pm.M));
  RBC_TRY(pm.vvC.resize(3 * pm.M));
  RBC_TRY(pm.vv.resize(3 * pm.G));
  // B-spline modulus factors, ModPME.F90:310-336
  double MP[64];
  int imin;
  h_bspline_func((double)p.P + 2.220446049250313e-16, p.P, &imin, MP);
  const int cnt[3] = {pm.Nxh, pm.Ny, pm.Nz};
  dbuf<double> *dst[3] = {&pm.bx, &pm.by, &pm.bz};
  for (int ii = 0; ii < 3; ii++) {
    std::vector<double> b(cnt[ii]);
    for (int k = 0; k < cnt[ii]; k++) {
      std::complex<double> s(0.0, 0.0);
      for (int m = 0; m <= p.P - 2; m++) {
        double ang = 2 * 3.14159265358979323846 * k * m / (double)p.Nb[ii];
        s += MP[m] * std::complex<double>(cos(ang), sin(ang));
      }
      double ang = 2 * 3.14159265358979323846 * k * (p.P - 1.) / (double)p.Nb[ii];
      std::complex<double> r = std::complex<double>(cos(ang), sin(ang)) / s;
      double a = std::abs(r);
      b[k] = a * a;
    }
    RBC_TRY(dst[ii]->resize(cnt[ii]));
    CUDA_TRY(cudaMemcpy(dst[ii]->p, b.data(), sizeof(double) * cnt[ii], cudaMemcpyHostToDevice));
  }
  return RBC3D_OK;
}

void pme_destroy(rbc3d_ctx *c) {
  Pme &pm = c->pme;
  for (int i = 0; i < 3; i++)
    if (pm.planF_ok[i]) {
      alloc_epoch()++;  // plan work areas are device memory a captured graph may refer to
      cufftDestroy(pm.planF[i]);
    }
  if (pm.planB_ok) {
    alloc_epoch()++;
    cufftDestroy(pm.planB);
  }
  if (pm.plan2F_ok) cufftDestroy(pm.plan2F);
  if (pm.plan2B_ok) cufftDestroy(pm.plan2B);
  for (int i = 0; i < 2; i++)
    if (pm.plan1_ok[i]) cufftDestroy(pm.plan1[i]);
  pm.sbuf.release(), pm.rbuf.release(), pm.Tz.release(), pm.Vz.release();
  pm.d_zoff.release(), pm.d_yoff.release(), pm.halo_tmp.release();
  pm.src.release();
  pm.srcC.release();
  pm.vvC.release();
  pm.vv.release();
  pm.bx.release();
  pm.by.release();
  pm.bz.release();
}

static int get_plan_fwd(rbc3d_ctx *c, int batch, cufftHandle *out) {
  Pme &pm = c->pme;
  const int slot = batch / 3 - 1;
  if (!pm.planF_ok[slot]) {
    int n[3] = {pm.Nz, pm.Ny, pm.Nx};
    CUFFT_TRY(cufftPlanMany(&pm.planF[slot], 3, n, nullptr, 1, (int)pm.G, nullptr, 1, (int)pm.M, CUFFT_D2Z, batch));
    pm.planF_ok[slot] = true;
  }
  *out = pm.planF[slot];
  CUFFT_TRY(cufftSetStream(pm.planF[slot], c->stream));  // the chain may run on either stream of the context
  return RBC3D_OK;
}

static int get_plan_bwd(rbc3d_ctx *c, cufftHandle *out) {
  Pme &pm = c->pme;
  if (!pm.planB_ok) {
    int n[3] = {pm.Nz, pm.Ny, pm.Nx};
    CUFFT_TRY(cufftPlanMany(&pm.planB, 3, n, nullptr, 1, (int)pm.M, nullptr, 1, (int)pm.G, CUFFT_Z2D, 3));
    pm.planB_ok = true;
  }
  *out = pm.planB;
  CUFFT_TRY(cufftSetStream(pm.planB, c->stream));
  return RBC3D_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Spreading (Distrib_Source, ModPME.F90:405-443).  One CTA per PME block of PME_BLK^3 mesh cells; the sources of
// the block (sorted by block) are accumulated into a shared-memory tile of (PME_BLK+P-1)^3 mesh points with
// plain read-modify-write (thread = one point of the P^3 support, one source at a time), then the tile is flushed
// to the mesh with native FP64 global reductions (halo points are shared with neighbouring blocks).
struct SpreadArgs {
  Params prm;
  int n;                 // number of points in the source list
  const int *start;      // block offsets
  const int *order;      // sorted position -> source index
  const double *x;       // SoA(3,n)
  const double *f;       // SoA(3,n) or null
  const double *g;       // SoA(3,n) or null
  const double *a3;      // SoA(3,n)
  const double *Bcell;   // per cell
  int npc;
  double c1, c2;
  int ncomp, comp0;      // components written: [comp0, comp0+ncomp)
  int nbx, nby, nbz;
  double *mesh;          // [9][G]
  size_t G;
  const CellList *list;  // host pointer: the PME list the arguments were taken from
};

template <int NCOMP>
__global__ void __launch_bounds__(512) k_spread(SpreadArgs a) {
  extern __shared__ double sm[];
  const int P = a.prm.P;
  const int T = PME_BLK + P - 1, T3 = T * T * T;
  double *tile = sm;                                   // [NCOMP][T3]
  double *sw = tile + (size_t)NCOMP * T3;              // [CHUNK][3][PMAX]
  double *sstr = sw + SPREAD_CHUNK * 3 * PME_PMAX;     // [CHUNK][NCOMP]
  int *srel = (int *)(sstr + SPREAD_CHUNK * NCOMP);    // [CHUNK][3]
  const int blk = blockIdx.x;
  const int sb = a.start[blk], se = a.start[blk + 1];
  if (sb == se) return;
  const int bx = blk % a.nbx, by = (blk / a.nbx) % a.nby, bz = blk / (a.nbx * a.nby);
  for (int i = threadIdx.x; i < NCOMP * T3; i += blockDim.x) tile[i] = 0.0;
  const int tid = threadIdx.x;
  const int i0 = tid % P, j0 = (tid / P) % P, k0 = tid / (P * P);
  const bool worker = k0 < P;
  for (int cb = sb; cb < se; cb += SPREAD_CHUNK) {
    const int nch = min(SPREAD_CHUNK, se - cb);
    __syncthreads();
    if (tid < nch * 3) {
      const int s = tid / 3, ax = tid - 3 * s;
      const int p = a.order[cb + s];
      const double u = __dmul_rn(a.x[(size_t)ax * a.n + p], a.prm.ih[ax]);  // ic = x*ih, ModPME.F90:420
      int imin;
      double w[PME_PMAX];
      bspline_func<PME_PMAX>(u, P, imin, w);
      const int mcell = imodulo(imin + (P - 1), a.prm.Nb[ax]);
      const int b = (ax == 0) ? bx : (ax == 1) ? by : bz;
      srel[s * 3 + ax] = mcell - b * PME_BLK;
      for (int q = 0; q < P; q++) sw[(s * 3 + ax) * PME_PMAX + q] = w[q];
    }
    for (int t = tid; t < nch * NCOMP; t += blockDim.x) {
      const int s = t / NCOMP, cc = t - s * NCOMP;
      const int p = a.order[cb + s];
      const size_t n = a.n;
      double v;
      int comp = a.comp0 + cc;
      if (comp < 3) {
        v = a.c1 * a.f[comp * n + p];
      } else {
        const double B = a.Bcell[p / a.npc];
        const double g0 = a.g[p], g1 = a.g[n + p], g2 = a.g[2 * n + p];
        const double n0 = a.a3[p] * B, n1 = a.a3[n + p] * B, n2 = a.a3[2 * n + p] * B;
        switch (comp) {
          case 3: v = g0 * n0; break;
          case 4: v = g1 * n1; break;
          case 5: v = g2 * n2; break;
          case 6: v = 0.5 * (g0 * n1 + g1 * n0); break;
          case 7: v = 0.5 * (g0 * n2 + g2 * n0); break;
          default: v = 0.5 * (g1 * n2 + g2 * n1); break;
        }
        v *= a.c2;
      }
      sstr[s * NCOMP + cc] = v;
    }
    __syncthreads();
    for (int s = 0; s < nch; s++) {
      if (worker) {
        const double w = sw[(s * 3 + 0) * PME_PMAX + i0] * sw[(s * 3 + 1) * PME_PMAX + j0] *
                         sw[(s * 3 + 2) * PME_PMAX + k0];
        const int addr = ((srel[s * 3 + 2] + k0) * T + (srel[s * 3 + 1] + j0)) * T + srel[s * 3 + 0] + i0;
#pragma unroll
        for (int cc = 0; cc < NCOMP; cc++) tile[cc * T3 + addr] += w * sstr[s * NCOMP + cc];
      }
      __syncthreads();
    }
  }
  // flush
  const int ox = bx * PME_BLK - (P - 1), oy = by * PME_BLK - (P - 1), oz = bz * PME_BLK - (P - 1);
  const int Nx = a.prm.Nb[0], Ny = a.prm.Nb[1], Nz = a.prm.Nb[2];
  for (int i = threadIdx.x; i < T3; i += blockDim.x) {
    const int lx = i % T, ly = (i / T) % T, lz = i / (T * T);
    const int gx = imodulo(ox + lx, Nx), gy = imodulo(oy + ly, Ny), gz = imodulo(oz + lz, Nz);
    const size_t gi = ((size_t)gz * Ny + gy) * Nx + gx;
#pragma unroll
    for (int cc = 0; cc < NCOMP; cc++) {
      const double v = tile[cc * T3 + i];
      if (v != 0.0) atomicAdd(a.mesh + (size_t)(a.comp0 + cc) * a.G + gi, v);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------
// P = 8 spreading, v2: register accumulation, no shared-memory read-modify-write and no barrier per source.
// One CTA per source block of 8 x 4 x 4 mesh cells; its sources touch a tile of 15 x 11 x 11 mesh points.  Thread =
// one (y, z) column of the tile, holding the 15 x-points x NCOMP components of that column in registers.  Per
// source: w_y w_z from zero-padded weight rows (no branch), then 8 x (1 + NCOMP) FMAs into the 8 x-points the
// source touches; the x offset is warp-uniform, so a switch selects one of 8 fully unrolled bodies (static register
// indices).  Weights / strengths of 32 sources at a time are prepared cooperatively in shared memory.  The tile is
// flushed component by component through a 15 x 11 x 11 shared buffer with coalesced FP64 reductions.
constexpr int SPR_TX = SPR_BX + 7, SPR_TY = SPR_BY + 7, SPR_TZ = SPR_BZ + 7;
constexpr int SPR_COLS = SPR_TY * SPR_TZ;  // 121
constexpr int SPR_THREADS = 128;
constexpr int SPR_PADW = 8 + 2 * 3 + 2;    // zero-padded weight row: index (t - rel + 3), t - rel in [-3, 10]

template <int NCOMP, int R>
__device__ __forceinline__ void spread_update(double (&acc)[SPR_TX][NCOMP], const double wyz,
                                              const double *__restrict__ wx, const double *__restrict__ st) {
  double sv[NCOMP];
#pragma unroll
  for (int cc = 0; cc < NCOMP; cc++) sv[cc] = st[cc];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const double w = wyz * wx[i];
#pragma unroll
    for (int cc = 0; cc < NCOMP; cc++) acc[R + i][cc] = fma(w, sv[cc], acc[R + i][cc]);
  }
}

template <int NCOMP>
__global__ void __launch_bounds__(SPR_THREADS) k_spread8(SpreadArgs a) {
  __shared__ __align__(16) double s_wx[SPREAD_CHUNK][8];
  __shared__ __align__(16) double s_wy[SPREAD_CHUNK][SPR_PADW];
  __shared__ __align__(16) double s_wz[SPREAD_CHUNK][SPR_PADW];
  __shared__ __align__(16) double s_str[SPREAD_CHUNK][NCOMP + (NCOMP & 1)];
  __shared__ int s_rel[SPREAD_CHUNK][4];
  extern __shared__ __align__(128) double s_flush[];  // [NCOMP][SPR_COLS][16] rows of the flush
  const int blk = blockIdx.x;
  const int sb = a.start[blk], se = a.start[blk + 1];
  if (sb == se) return;
  const int bx = blk % a.nbx, by = (blk / a.nbx) % a.nby, bz = blk / (a.nbx * a.nby);
  const int tid = threadIdx.x;
  const bool col = tid < SPR_COLS;
  const int ty = tid % SPR_TY, tz = col ? tid / SPR_TY : 0;
  double acc[SPR_TX][NCOMP];
#pragma unroll
  for (int i = 0; i < SPR_TX; i++)
#pragma unroll
    for (int cc = 0; cc < NCOMP; cc++) acc[i][cc] = 0.0;
  for (int i = tid; i < SPREAD_CHUNK * SPR_PADW; i += SPR_THREADS) {
    (&s_wy[0][0])[i] = 0.0;
    (&s_wz[0][0])[i] = 0.0;
  }
  for (int cb = sb; cb < se; cb += SPREAD_CHUNK) {
    const int nch = min(SPREAD_CHUNK, se - cb);
    __syncthreads();
    if (tid < nch * 3) {
      const int s = tid / 3, ax = tid - 3 * s;
      const int p = a.order[cb + s];
      const double u = __dmul_rn(a.x[(size_t)ax * a.n + p], a.prm.ih[ax]);  // ic = x*ih, ModPME.F90:420
      int imin;
      double w[PME_PMAX];
      bspline_func<PME_PMAX>(u, 8, imin, w);
      const int mcell = imodulo(imin + 7, a.prm.Nb[ax]);
      const int b = (ax == 0) ? bx * SPR_BX : (ax == 1) ? by * SPR_BY : bz * SPR_BZ;
      s_rel[s][ax] = mcell - b;
      double *dst = (ax == 0) ? &s_wx[s][0] : (ax == 1) ? &s_wy[s][3] : &s_wz[s][3];
#pragma unroll
      for (int q = 0; q < 8; q++) dst[q] = w[q];
    }
    for (int t = tid; t < nch * NCOMP; t += SPR_THREADS) {
      const int s = t / NCOMP, cc = t - s * NCOMP;
      const int p = a.order[cb + s];
      const size_t n = a.n;
      double v;
      const int comp = a.comp0 + cc;
      if (comp < 3) {
        v = a.c1 * a.f[comp * n + p];
      } else {
        const double B = a.Bcell[p / a.npc];
        const double g0 = a.g[p], g1 = a.g[n + p], g2 = a.g[2 * n + p];
        const double n0 = a.a3[p] * B, n1 = a.a3[n + p] * B, n2 = a.a3[2 * n + p] * B;
        switch (comp) {
          case 3: v = g0 * n0; break;
          case 4: v = g1 * n1; break;
          case 5: v = g2 * n2; break;
          case 6: v = 0.5 * (g0 * n1 + g1 * n0); break;
          case 7: v = 0.5 * (g0 * n2 + g2 * n0); break;
          default: v = 0.5 * (g1 * n2 + g2 * n1); break;
        }
        v *= a.c2;
      }
      s_str[s][cc] = v;
    }
    __syncthreads();
    if (col) {
      for (int s = 0; s < nch; s++) {
        const int rx = s_rel[s][0], ry = s_rel[s][1], rz = s_rel[s][2];
        const double wyz = s_wy[s][ty - ry + 3] * s_wz[s][tz - rz + 3];
        const double *wx = s_wx[s], *st = s_str[s];
        switch (rx) {
          case 0: spread_update<NCOMP, 0>(acc, wyz, wx, st); break;
          case 1: spread_update<NCOMP, 1>(acc, wyz, wx, st); break;
          case 2: spread_update<NCOMP, 2>(acc, wyz, wx, st); break;
          case 3: spread_update<NCOMP, 3>(acc, wyz, wx, st); break;
          case 4: spread_update<NCOMP, 4>(acc, wyz, wx, st); break;
          case 5: spread_update<NCOMP, 5>(acc, wyz, wx, st); break;
          case 6: spread_update<NCOMP, 6>(acc, wyz, wx, st); break;
          default: spread_update<NCOMP, 7>(acc, wyz, wx, st); break;
        }
      }
    }
  }
  // flush: every thread owns the x-row of its (y,z) column in registers.  The row goes to shared memory as 16
  // doubles [x0-8 .. x0+7] (first one a zero pad: 128-byte rows, 64-byte aligned in the mesh) and is added to the mesh
  // by ONE bulk reduction per row and component (cp.reduce.async.bulk ... add.f64: the additions happen in L2 and
  // cost the SM one instruction per row instead of 15 lane-wise reductions with their index arithmetic).
  if (col) {
    const int Nx = a.prm.Nb[0], Ny = a.prm.Nb[1], Nz = a.prm.Nb[2];
    const int gy = imodulo(by * SPR_BY - 7 + ty, Ny), gz = imodulo(bz * SPR_BZ - 7 + tz, Nz);
    const int gx0 = bx * SPR_BX - 8;  // multiple of 8; only the first block of a row wraps (gx0 = -8)
    double *srow = s_flush + (size_t)tid * 16;
    bool issued = false;
#pragma unroll
    for (int cc = 0; cc < NCOMP; cc++) {
      bool any = false;
#pragma unroll
      for (int i = 0; i < SPR_TX; i++) any = any || (acc[i][cc] != 0.0);
      if (!any) continue;
      double *r = srow + (size_t)cc * SPR_COLS * 16;
      r[0] = 0.0;
#pragma unroll
      for (int i = 0; i < SPR_TX; i++) r[1 + i] = acc[i][cc];
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      double *mrow = a.mesh + (size_t)(a.comp0 + cc) * a.G + ((size_t)gz * Ny + gy) * Nx;
      const unsigned sa = (unsigned)__cvta_generic_to_shared(r);
      if (gx0 >= 0 && gx0 + 16 <= Nx) {
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 128;" ::"l"(mrow + gx0), "r"(sa)
                     : "memory");
      } else if (gx0 >= 0 || Nx < 16) {  // mesh width not a multiple of 8: the last block wraps on the right
#pragma unroll
        for (int i = 0; i < SPR_TX; i++)
          if (acc[i][cc] != 0.0) atomicAdd(mrow + imodulo(gx0 + 1 + i, Nx), acc[i][cc]);
        continue;
      } else {  // periodic wrap of the left halo: [Nx-8, Nx) and [0, 8)
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 64;" ::"l"(mrow + Nx - 8), "r"(sa)
                     : "memory");
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 64;" ::"l"(mrow), "r"(sa + 64)
                     : "memory");
      }
      issued = true;
    }
    if (issued) {
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared rows must outlive the reads
    }
  }
}


// ---------------------------------------------------------------------------------------------------------
// Record staging of the P = 8 walk kernels.  A warp consumes the sorted points of its column in rounds of 8; the
// records of round r + 1 (weights, mesh cell x, point index or strengths: geometry-time / pre-pass products read
// exactly once) are copied global -> shared with cp.async while round r is evaluated, so no global-memory latency
// sits on the per-point critical path.
constexpr int WK_ROUND = 8, WK_REC = 32;  // shared record: 26 doubles of the list record, [26..29] aux, pad

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// stage the records of sorted points [s0, s0 + 8) (clamped to last): 13 16-byte chunks of the list record each, plus
// either a 32-byte strength record (aux32, spreading) or the point index (order, interpolation)
__device__ __forceinline__ void wk_stage(double *buf, int lane, int s0, int last, const double *__restrict__ wrec,
                                         const double *__restrict__ aux32, const int *__restrict__ order) {
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int ch = lane + 32 * k;
    if (ch < WK_ROUND * 13) {
      const int slot = ch / 13, part = ch - 13 * slot;
      const int s = min(s0 + slot, last);
      cp_async16(buf + slot * WK_REC + 2 * part, wrec + (size_t)s * PME_WREC + 2 * part);
    }
  }
  if (aux32) {
    if (lane < 2 * WK_ROUND) {
      const int slot = lane >> 1, part = lane & 1;
      const int s = min(s0 + slot, last);
      cp_async16(buf + slot * WK_REC + 26 + 2 * part, aux32 + (size_t)s * 4 + 2 * part);
    }
  } else if (lane < WK_ROUND) {
    const int s = min(s0 + lane, last);
    cp_async4(buf + lane * WK_REC + 26, order + s);
  }
  cp_async_commit();
}

// ---------------------------------------------------------------------------------------------------------
// P = 8 spreading, v3: register-ring pencil walk.  One warp per (x run of 8 mesh cells, y cell, component triple);
// its sources are sorted by z cell.  Every source of the pencil touches the same 15 x 8 (x, y) footprint, so no lane
// ever multiplies by a padded zero weight (the block kernel k_spread8 wastes 47 % of its columns).  Lane = (y offset,
// z-slot pair): a ring of the last 8 z planes, two slots per lane, 15 x-points x 3 components each = 90 accumulators in
// registers.  Per source and lane: 8 multiplies for w_y w_z (strength) and 48 FMAs with w_x; the x offset of the
// source's cell inside the run is warp-uniform and selects one of 8 unrolled bodies.  Moving up in z retires ring
// slots: the 8 lanes that own the retiring plane write their 15-point rows to a double-buffered shared staging
// area and add them to the mesh with one 128-byte bulk reduction per row and component (cp.reduce.async.bulk
// add.f64, performed in L2); every row is written by the lane that issues its reduction, so no barrier is needed.
// Weights are geometry-time records, strengths come from a pre-pass in sorted order (k_spread_strength); both are
// staged one round ahead by wk_stage.
constexpr int SW_WARPS = 2, SW_TX = SW_XR + 7;
constexpr int SW_STG = 3 * 16 + 2;  // per-lane staging stride: 400 B, lanes 16 B apart in the banks

struct SpreadWArgs {
  Params prm;
  int n;
  const int *start;          // pencil list, key = cz + Nz * (xrun + nxr * cy)
  const double *wrec;        // [sorted source][PME_WREC]
  const double *str;         // [sorted source][npass][4]
  int comp0, npass, nxr;     // components [comp0, comp0 + 3 npass)
  double *mesh;              // [9][G]
  size_t G;
};

// strengths of the sorted sources: c1 f (pass 0 of the single layer) or c2 sym(g (x) a3 B) (diagonal, off-diagonal)
struct StrengthArgs {
  int ns, n, npc, dl;
  const int *order;
  const double *f, *g, *a3, *Bcell;
  double c1, c2;
  double *str;
};
__global__ void k_spread_strength(StrengthArgs a) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.ns) return;
  const int p = a.order[s];
  const size_t n = a.n;
  if (!a.dl) {
    double4 v = make_double4(a.c1 * a.f[p], a.c1 * a.f[n + p], a.c1 * a.f[2 * n + p], 0.0);
    reinterpret_cast<double4 *>(a.str)[s] = v;
  } else {
    const double B = a.Bcell[p / a.npc];
    const double g0 = a.g[p], g1 = a.g[n + p], g2 = a.g[2 * n + p];
    const double n0 = a.a3[p] * B, n1 = a.a3[n + p] * B, n2 = a.a3[2 * n + p] * B;
    double4 d = make_double4(a.c2 * (g0 * n0), a.c2 * (g1 * n1), a.c2 * (g2 * n2), 0.0);
    double4 o = make_double4(a.c2 * (0.5 * (g0 * n1 + g1 * n0)), a.c2 * (0.5 * (g0 * n2 + g2 * n0)),
                             a.c2 * (0.5 * (g1 * n2 + g2 * n1)), 0.0);
    reinterpret_cast<double4 *>(a.str)[2 * (size_t)s] = d;
    reinterpret_cast<double4 *>(a.str)[2 * (size_t)s + 1] = o;
  }
}

template <int RX>
__device__ __forceinline__ void sw_update(double (&A)[2][SW_TX][3], const double *__restrict__ w,
                                          const double (&s0)[3], const double (&s1)[3]) {
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const double2 wx = reinterpret_cast<const double2 *>(w)[k];
#pragma unroll
    for (int cc = 0; cc < 3; cc++) {
      A[0][RX + 2 * k][cc] = fma(wx.x, s0[cc], A[0][RX + 2 * k][cc]);
      A[1][RX + 2 * k][cc] = fma(wx.x, s1[cc], A[1][RX + 2 * k][cc]);
      A[0][RX + 2 * k + 1][cc] = fma(wx.y, s0[cc], A[0][RX + 2 * k + 1][cc]);
      A[1][RX + 2 * k + 1][cc] = fma(wx.y, s1[cc], A[1][RX + 2 * k + 1][cc]);
    }
  }
}

// add the 15-point rows of ring column COL (3 components) of this lane to mesh rows (gy, gz) and clear them.  The rows
// go through this lane's private staging area [3][16] (first entry a zero pad: 128-byte rows, 64-byte aligned in the
// mesh) and leave as one bulk reduction per row; meshes narrower than 16 or a last run that wraps on the right take a
// rolled loop of scalar reductions from the same staging rows.
template <int COL>
__device__ __forceinline__ void sw_flush(double (&A)[2][SW_TX][3], double *stage, const SpreadWArgs &a, int comp_first,
                                         int gx0, int gy, int gz) {
  const int Nx = a.prm.Nb[0], Ny = a.prm.Nb[1];
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the previous flush of this lane has left the stage
#pragma unroll
  for (int cc = 0; cc < 3; cc++) {
    double *r = stage + cc * 16;
    r[0] = 0.0;
#pragma unroll
    for (int i = 0; i < SW_TX; i++) {
      r[1 + i] = A[COL][i][cc];
      A[COL][i][cc] = 0.0;
    }
  }
  double *mrow = a.mesh + (size_t)comp_first * a.G + ((size_t)gz * Ny + gy) * Nx;
  const unsigned sa = (unsigned)__cvta_generic_to_shared(stage);
  if ((gx0 >= 0 && gx0 + 16 <= Nx) || (gx0 < 0 && Nx >= 16)) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll 1
    for (int cc = 0; cc < 3; cc++) {
      double *m = mrow + (size_t)cc * a.G;
      const unsigned sr = sa + cc * 128;
      if (gx0 >= 0) {
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 128;" ::"l"(m + gx0), "r"(sr)
                     : "memory");
      } else {  // periodic wrap of the left halo: [Nx-8, Nx) and [0, 8)
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 64;" ::"l"(m + Nx - 8), "r"(sr)
                     : "memory");
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 64;" ::"l"(m), "r"(sr + 64)
                     : "memory");
      }
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  } else {
#pragma unroll 1
    for (int k = 0; k < 3 * 16; k++) {
      const int cc = k >> 4, i = (k & 15) - 1;
      const double v = stage[k];
      if (i >= 0 && v != 0.0) atomicAdd(mrow + (size_t)cc * a.G + imodulo(gx0 + 1 + i, Nx), v);
    }
  }
}

__global__ void __launch_bounds__(SW_WARPS * 32, 4) k_spread_walk(SpreadWArgs a) {
  __shared__ __align__(16) double s_rec[SW_WARPS][2][WK_ROUND * WK_REC];
  __shared__ __align__(128) double s_fl[SW_WARPS][32][SW_STG];  // flush staging, private to a lane
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Ny = a.prm.Nb[1], Nz = a.prm.Nb[2];
  const int task = blockIdx.x * SW_WARPS + warp;
  const int pencil = task / a.npass, pass = task - pencil * a.npass;
  if (pencil >= a.nxr * Ny) return;
  const int *st = a.start + (size_t)pencil * Nz;
  const int pos0 = st[0], end = st[Nz];
  if (pos0 == end) return;
  const double *aux = a.str + 4 * pass;  // record of point s at aux + s * 4 * npass
  // (wk_stage addresses aux32 + s * 4: fold npass into the pointer arithmetic below)
  const int xr = pencil % a.nxr, cy = pencil / a.nxr;
  const int jy = lane & 7, zq = lane >> 3;
  const int gy = imodulo(cy - 7 + jy, Ny);
  const int gx0 = xr * SW_XR - 8;
  const int comp_first = a.comp0 + 3 * pass;
  auto stage = [&](double *buf, int s0) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int ch = lane + 32 * k;
      if (ch < WK_ROUND * 13) {
        const int slot = ch / 13, part = ch - 13 * slot;
        const int s = min(s0 + slot, end - 1);
        cp_async16(buf + slot * WK_REC + 2 * part, a.wrec + (size_t)s * PME_WREC + 2 * part);
      }
    }
    if (lane < 2 * WK_ROUND) {
      const int slot = lane >> 1, part = lane & 1;
      const int s = min(s0 + slot, end - 1);
      cp_async16(buf + slot * WK_REC + 26 + 2 * part, aux + (size_t)s * 4 * a.npass + 2 * part);
    }
    cp_async_commit();
  };
  stage(s_rec[warp][0], pos0);
  double A[2][SW_TX][3];
#pragma unroll
  for (int col = 0; col < 2; col++)
#pragma unroll
    for (int i = 0; i < SW_TX; i++)
#pragma unroll
      for (int cc = 0; cc < 3; cc++) A[col][i][cc] = 0.0;
  int pos = pos0, cz = 0, sprev = -64, cur_round = -1;
  double *stg = &s_fl[warp][lane][0];
  int zwin = 0;
  int e = (lane < Nz) ? st[lane + 1] : end + 1;
  for (;;) {
    int cnt = 0;
    if (pos < end) {
      for (;;) {  // next occupied z cell of the pencil
        const unsigned m = __ballot_sync(FULL_MASK, e > pos);
        if (m) {
          const int f = __ffs(m) - 1;
          cnt = __shfl_sync(FULL_MASK, e, f) - pos;
          cz = zwin + f;
          break;
        }
        zwin += 32;
        e = (zwin + lane < Nz) ? st[zwin + lane + 1] : end + 1;
      }
    } else {
      cz = sprev + 8;  // past the last source: everything retires
    }
    if (sprev >= 0) {  // retire this lane's ring slots (zq, zq + 4) whose planes lie in [sprev - 7, cz - 8]
      const int u0 = sprev - ((sprev - zq) & 7), u1 = sprev - ((sprev - zq - 4) & 7);
      if (u0 <= cz - 8) sw_flush<0>(A, stg, a, comp_first, gx0, gy, imodulo(u0, Nz));
      if (u1 <= cz - 8) sw_flush<1>(A, stg, a, comp_first, gx0, gy, imodulo(u1, Nz));
    }
    if (pos >= end) break;
    for (int t = 0; t < cnt; t++) {
      const int rel = pos + t - pos0, r = rel >> 3;
      if (r != cur_round) {  // warp-uniform: enter round r (staged one round ahead), prefetch round r + 1
        cp_async_wait_all();
        __syncwarp();
        cur_round = r;
        const int s1 = pos0 + 8 * (r + 1);
        if (s1 < end) stage(s_rec[warp][(r + 1) & 1], s1);
      }
      const double *w = s_rec[warp][r & 1] + (rel & 7) * WK_REC;
      const int rx = __double2loint(w[24]) - xr * SW_XR;
      const double wy = w[8 + jy];
      const double wyz0 = wy * w[16 + zq], wyz1 = wy * w[20 + zq];  // z weights are stored in ring order
      double s0[3], s1v[3];
#pragma unroll
      for (int cc = 0; cc < 3; cc++) {
        const double sv = w[26 + cc];
        s0[cc] = wyz0 * sv;
        s1v[cc] = wyz1 * sv;
      }
      switch (rx) {
        case 0: sw_update<0>(A, w, s0, s1v); break;
        case 1: sw_update<1>(A, w, s0, s1v); break;
        case 2: sw_update<2>(A, w, s0, s1v); break;
        case 3: sw_update<3>(A, w, s0, s1v); break;
        case 4: sw_update<4>(A, w, s0, s1v); break;
        case 5: sw_update<5>(A, w, s0, s1v); break;
        case 6: sw_update<6>(A, w, s0, s1v); break;
        default: sw_update<7>(A, w, s0, s1v); break;
      }
    }
    sprev = cz;
    pos += cnt;
  }
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared rows must outlive the reads
}

// launch the spreading kernel(s) for one source list (cells or wall centroids); the meshes are accumulated into
static int spread_launch(rbc3d_ctx *c, SpreadArgs a, bool sl, bool dl) {
  Pme &pm = c->pme;
  struct { bool flag_sl, flag_dl; } pmf = {sl, dl};
    const CellList &L = *a.list;
    a.nbx = L.nsblk[0];
    a.nby = L.nsblk[1];
    a.nbz = L.nsblk[2];
    a.mesh = pm.src.p;
    a.G = pm.G;
    const int nblocks = a.nbx * a.nby * a.nbz;
    if (L.swalk) {
      const int ns = L.n_sorted;
      if (ns > 0) {
        SpreadWArgs w;
        w.prm = a.prm, w.n = a.n, w.start = a.start, w.wrec = L.w.p;
        w.nxr = L.nsblk[0], w.mesh = pm.src.p, w.G = pm.G;
        for (int which = 0; which < 2; which++) {
          if (which == 0 ? !pmf.flag_sl : !pmf.flag_dl) continue;
          w.comp0 = which == 0 ? 0 : 3;
          w.npass = which == 0 ? 1 : 2;
          RBC_TRY(pm.str.resize((size_t)ns * 4 * w.npass));
          StrengthArgs sa;
          sa.ns = ns, sa.n = a.n, sa.npc = a.npc, sa.dl = which, sa.order = a.order;
          sa.f = a.f, sa.g = a.g, sa.a3 = a.a3, sa.Bcell = a.Bcell, sa.c1 = a.c1, sa.c2 = a.c2, sa.str = pm.str.p;
          k_spread_strength<<<(ns + 255) / 256, 256, 0, c->stream>>>(sa);
          w.str = pm.str.p;
          const int ntask = L.nsblk[0] * pm.Ny * w.npass;
          k_spread_walk<<<(ntask + SW_WARPS - 1) / SW_WARPS, SW_WARPS * 32, 0, c->stream>>>(w);
          c->launches += 2;
        }
        KERNEL_CHECK();
      }
    } else if (c->prm.P == 8) {
      if (pmf.flag_sl) {
        a.comp0 = 0;
        a.ncomp = 3;
        CUDA_TRY(cudaFuncSetAttribute(k_spread8<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * SPR_COLS * 128));
        k_spread8<3><<<nblocks, SPR_THREADS, 3 * SPR_COLS * 128, c->stream>>>(a);
        c->launches++;
      }
      if (pmf.flag_dl) {
        a.comp0 = 3;
        a.ncomp = 6;
        CUDA_TRY(cudaFuncSetAttribute(k_spread8<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * SPR_COLS * 128));
        k_spread8<6><<<nblocks, SPR_THREADS, 6 * SPR_COLS * 128, c->stream>>>(a);
        c->launches++;
      }
      KERNEL_CHECK();
    } else {
    const int T = PME_BLK + c->prm.P - 1;
    auto smem = [&](int nc) {
      return sizeof(double) * ((size_t)nc * T * T * T + SPREAD_CHUNK * 3 * PME_PMAX + SPREAD_CHUNK * nc) +
             sizeof(int) * SPREAD_CHUNK * 3;
    };
    if (pmf.flag_sl && pmf.flag_dl) {
      a.comp0 = 0;
      a.ncomp = 9;
      CUDA_TRY(cudaFuncSetAttribute(k_spread<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem(9)));
      k_spread<9><<<nblocks, 512, smem(9), c->stream>>>(a);
    } else if (pmf.flag_sl) {
      a.comp0 = 0;
      a.ncomp = 3;
      CUDA_TRY(cudaFuncSetAttribute(k_spread<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem(3)));
      k_spread<3><<<nblocks, 512, smem(3), c->stream>>>(a);
    } else {
      a.comp0 = 3;
      a.ncomp = 6;
      CUDA_TRY(cudaFuncSetAttribute(k_spread<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem(6)));
      k_spread<6><<<nblocks, 512, smem(6), c->stream>>>(a);
    }
    KERNEL_CHECK();
    c->launches++;
    }
  return RBC3D_OK;
}

int pme_spread(rbc3d_ctx *c, double c1, double c2, bool use_cells, bool use_walls) {
  Pme &pm = c->pme;
  Cells &C = c->cells;
  pm.flag_sl = fabs(c1) > 1.e-10;  // ModPME.F90:71-72
  pm.flag_dl = fabs(c2) > 1.e-10;
  pm.transformed = false;
  if (pm.flag_sl) CUDA_TRY(cudaMemsetAsync(pm.src.p, 0, sizeof(double) * 3 * pm.G, c->stream));
  if (pm.flag_dl) CUDA_TRY(cudaMemsetAsync(pm.src.p + 3 * pm.G, 0, sizeof(double) * 6 * pm.G, c->stream));
  if (use_cells && C.Np > 0 && (pm.flag_sl || pm.flag_dl)) {
    if (pm.flag_sl && !C.f_set) {
      set_error("PME_Distrib_Source: c1 != 0 but no single-layer density set");
      return RBC3D_ESTATE;
    }
    if (pm.flag_dl && !C.g_set) {
      set_error("PME_Distrib_Source: c2 != 0 but no double-layer density set");
      return RBC3D_ESTATE;
    }
    SpreadArgs a;
    a.prm = c->prm;
    a.n = C.Np;
    a.list = &C.pl;
    a.start = C.pl.start.p;
    a.order = C.pl.order.p;
    a.x = C.x.p;
    a.f = C.f.p;
    a.g = C.g.p;
    a.a3 = C.a3.p;
    a.Bcell = C.B.p;
    a.npc = C.npc;
    a.c1 = c1;
    a.c2 = c2;
    RBC_TRY(spread_launch(c, a, pm.flag_sl, pm.flag_dl));
  }
  // wall sources: element centroids with THRD*sum(fele)*area, single layer only (ModPME.F90:105-129)
  Walls &W = c->walls;
  if (use_walls && W.NE > 0 && pm.flag_sl) {
    if (!W.f_set) {
      set_error("PME_Distrib_Source: wall tractions not set");
      return RBC3D_ESTATE;
    }
    SpreadArgs a;
    a.prm = c->prm;
    a.n = W.NE;
    a.list = &W.pl;
    a.start = W.pl.start.p;
    a.order = W.pl.order.p;
    a.x = W.xc.p;
    a.f = W.ft.p;
    a.g = nullptr;
    a.a3 = nullptr;
    a.Bcell = nullptr;
    a.npc = 1;
    a.c1 = c1;
    a.c2 = 0.0;
    RBC_TRY(spread_launch(c, a, true, false));
  }
  if (c->prm.nranks > 1 && !pm.slab && (pm.flag_sl || pm.flag_dl)) {
    // (slab decomposition off) every rank spread its block of cells: sum the meshes over the ranks
    double *base = pm.src.p + (pm.flag_sl ? 0 : 3 * pm.G);
    const size_t ncomp = (pm.flag_sl ? 3 : 0) + (pm.flag_dl ? 6 : 0);
    RBC_TRY(comm_allreduce_sum(c, base, ncomp * pm.G));
  }
  pm.distributed = true;
  return RBC3D_OK;
}

// ---------------------------------------------------------------------------------------------------------
// k-space multiply (ModPME.F90:165-210), fused with the B-spline modulus and the Hermitian symmetrisation of the
// x = 0 and x = Nx/2 planes.  One thread per retained mode.
struct KArgs {
  Params prm;
  int Nx, Ny, Nz, Nxh;
  size_t M;
  const cufftDoubleComplex *srcC;  // [9][M]
  cufftDoubleComplex *vvC;         // [3][M]
  const double *bx, *by, *bz;
  int sl, dl;
  double vol;
  // element (i, j, k) of a component sits at i*sx + (j - j0)*sy + k*sz (components M apart): [Nz][Ny][Nxh] for the
  // single-rank transform, [y rows of the rank][Nxh][Nz] (z fastest) on the y-slabs of the decomposed transform
  long long sx, sy, sz;
  int j0;
};

template <bool SL, bool DL>
__device__ __forceinline__ void kspace_mode(const KArgs &a, int i, int j, int k, double vr[3], double vi[3]) {
  vr[0] = vr[1] = vr[2] = vi[0] = vi[1] = vi[2] = 0.0;
  if (i == 0 && j == 0 && k == 0) return;
  const size_t idx = (size_t)i * a.sx + (size_t)(j - a.j0) * a.sy + (size_t)k * a.sz;
  const double q0 = (double)i * a.prm.iLb[0];
  const double q1 = (double)(j < a.Ny / 2 ? j : j - a.Ny) * a.prm.iLb[1];
  const double q2 = (double)(k <= a.Nz / 2 ? k : k - a.Nz) * a.prm.iLb[2];
  const double alpha = a.prm.alpha;
  const double qq = q0 * q0 + q1 * q1 + q2 * q2;
  const double q2t = RBC_PI * alpha * qq;
  const double e = exp(-q2t);
  const double phi0 = e / q2t;
  const double phi1 = (e + phi0) / q2t;
  const double q[3] = {q0, q1, q2};
  if (SL) {
    double fr[3], fi[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
      const cufftDoubleComplex z = a.srcC[(size_t)d * a.M + idx];
      fr[d] = z.x;
      fi[d] = z.y;
    }
    // 2 alpha/V phi1 (q2t F - qt (qt.F)), qt = sqrt(pi alpha) q  => qt(qt.F) = pi alpha q (q.F)
    const double pa = RBC_PI * alpha;
    const double dr = q0 * fr[0] + q1 * fr[1] + q2 * fr[2], di = q0 * fi[0] + q1 * fi[1] + q2 * fi[2];
    const double cf = 2.0 * alpha / a.vol * phi1;
#pragma unroll
    for (int d = 0; d < 3; d++) {
      vr[d] += cf * (q2t * fr[d] - pa * q[d] * dr);
      vi[d] += cf * (q2t * fi[d] - pa * q[d] * di);
    }
  }
  if (DL) {
    double sr[6], si[6];
#pragma unroll
    for (int d = 0; d < 6; d++) {
      const cufftDoubleComplex z = a.srcC[(size_t)(3 + d) * a.M + idx];
      sr[d] = z.x;
      si[d] = z.y;
    }
    // S = [[0,3,4],[3,1,5],[4,5,2]]
    const double trr = sr[0] + sr[1] + sr[2], tri = si[0] + si[1] + si[2];
    const double Sqr[3] = {sr[0] * q0 + sr[3] * q1 + sr[4] * q2, sr[3] * q0 + sr[1] * q1 + sr[5] * q2,
                           sr[4] * q0 + sr[5] * q1 + sr[2] * q2};
    const double Sqi[3] = {si[0] * q0 + si[3] * q1 + si[4] * q2, si[3] * q0 + si[1] * q1 + si[5] * q2,
                           si[4] * q0 + si[5] * q1 + si[2] * q2};
    const double qSqr = q0 * Sqr[0] + q1 * Sqr[1] + q2 * Sqr[2];
    const double qSqi = q0 * Sqi[0] + q1 * Sqi[1] + q2 * Sqi[2];
    const double ca = 4.0 * RBC_PI * alpha / a.vol * phi0;
    const double cb = 8.0 * RBC_PI * RBC_PI * alpha * alpha / a.vol * phi1;
#pragma unroll
    for (int d = 0; d < 3; d++) {
      // t = i*[ca (q tr + 2 S q) - cb (q.Sq) q];  V -= t ;  i*(x+iy) = -y + ix
      const double wr = ca * (q[d] * trr + 2.0 * Sqr[d]) - cb * qSqr * q[d];
      const double wi = ca * (q[d] * tri + 2.0 * Sqi[d]) - cb * qSqi * q[d];
      vr[d] -= -wi;
      vi[d] -= wr;
    }
  }
  const double b = a.bx[i] * a.by[j] * a.bz[k];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    vr[d] *= b;
    vi[d] *= b;
  }
}

template <bool SL, bool DL>
__global__ void __launch_bounds__(256) k_kspace(KArgs a) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.M) return;
  const int i = (int)(idx % a.Nxh);
  const int j = (int)((idx / a.Nxh) % a.Ny);
  const int k = (int)(idx / ((size_t)a.Nxh * a.Ny));
  double vr[3], vi[3];
  kspace_mode<SL, DL>(a, i, j, k, vr, vi);
  if (i == 0 || i == a.Nx / 2) {
    double wr[3], wi[3];
    kspace_mode<SL, DL>(a, i, (a.Ny - j) % a.Ny, (a.Nz - k) % a.Nz, wr, wi);
#pragma unroll
    for (int d = 0; d < 3; d++) {
      vr[d] = 0.5 * (vr[d] + wr[d]);
      vi[d] = 0.5 * (vi[d] - wi[d]);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; d++) a.vvC[(size_t)d * a.M + idx] = make_cuDoubleComplex(vr[d], vi[d]);
}

int pme_transform(rbc3d_ctx *c) {
  Pme &pm = c->pme;
  if (!pm.distributed) {
    set_error("PME_Transform called before PME_Distrib_Source");
    return RBC3D_ESTATE;
  }
  const Params &p = c->prm;
  if (!pm.flag_sl && !pm.flag_dl) {
    CUDA_TRY(cudaMemsetAsync(pm.vv.p, 0, sizeof(double) * 3 * pm.G, c->stream));
    pm.transformed = true;
    return RBC3D_OK;
  }
  if (pm.slab) return pme_transform_slab(c);
  t_begin(c, RBC3D_T_FFT);
  cufftHandle plan;
  if (pm.flag_sl && pm.flag_dl) {
    RBC_TRY(get_plan_fwd(c, 9, &plan));
    CUFFT_TRY(cufftExecD2Z(plan, pm.src.p, pm.srcC.p));
  } else if (pm.flag_sl) {
    RBC_TRY(get_plan_fwd(c, 3, &plan));
    CUFFT_TRY(cufftExecD2Z(plan, pm.src.p, pm.srcC.p));
  } else {
    RBC_TRY(get_plan_fwd(c, 6, &plan));
    CUFFT_TRY(cufftExecD2Z(plan, pm.src.p + 3 * pm.G, pm.srcC.p + 3 * pm.M));
  }
  t_end(c, RBC3D_T_FFT);
  t_begin(c, RBC3D_T_KSPACE);
  KArgs a;
  a.prm = p;
  a.Nx = pm.Nx;
  a.Ny = pm.Ny;
  a.Nz = pm.Nz;
  a.Nxh = pm.Nxh;
  a.M = pm.M;
  a.srcC = pm.srcC.p;
  a.vvC = pm.vvC.p;
  a.bx = pm.bx.p;
  a.by = pm.by.p;
  a.bz = pm.bz.p;
  a.vol = p.Lb[0] * p.Lb[1] * p.Lb[2];
  a.sx = 1, a.sy = pm.Nxh, a.sz = (long long)pm.Ny * pm.Nxh, a.j0 = 0;
  const int grid = (int)((pm.M + 255) / 256);
  if (pm.flag_sl && pm.flag_dl)
    k_kspace<true, true><<<grid, 256, 0, c->stream>>>(a);
  else if (pm.flag_sl)
    k_kspace<true, false><<<grid, 256, 0, c->stream>>>(a);
  else
    k_kspace<false, true><<<grid, 256, 0, c->stream>>>(a);
  KERNEL_CHECK();
  t_end(c, RBC3D_T_KSPACE);
  cufftHandle planb;
  RBC_TRY(get_plan_bwd(c, &planb));
  t_begin(c, RBC3D_T_FFT_INV);
  CUFFT_TRY(cufftExecZ2D(planb, pm.vvC.p, pm.vv.p));
  t_end(c, RBC3D_T_FFT_INV);
  c->launches += 3;
  pm.transformed = true;
  return RBC3D_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Interpolation (Interp_Vel, ModPME.F90:450-489): one CTA per PME block of targets, the (PME_BLK+P-1)^3 x 3
// velocity tile is staged in shared memory (coalesced rows), one warp per target gathers its P^3 support.
struct InterpArgs {
  Params prm;
  int n;
  const int *start, *order;
  const double *x;   // targets SoA(3,n)
  const double *vv;  // [3][G]
  size_t G;
  int nbx, nby, nbz;
  double *acc;       // SoA(3,n)
};

constexpr int INTERP_WARPS = 8;

__global__ void __launch_bounds__(INTERP_WARPS * 32) k_interp(InterpArgs a) {
  extern __shared__ double sm[];
  const int P = a.prm.P;
  const int T = PME_BLK + P - 1, T3 = T * T * T;
  double *tile = sm;            // [3][T3]
  double *sw = tile + 3 * T3;   // [WARPS][3][PMAX]
  const int blk = blockIdx.x;
  const int sb = a.start[blk], se = a.start[blk + 1];
  if (sb == se) return;
  const int bx = blk % a.nbx, by = (blk / a.nbx) % a.nby, bz = blk / (a.nbx * a.nby);
  const int ox = bx * PME_BLK - (P - 1), oy = by * PME_BLK - (P - 1), oz = bz * PME_BLK - (P - 1);
  const int Nx = a.prm.Nb[0], Ny = a.prm.Nb[1], Nz = a.prm.Nb[2];
  for (int i = threadIdx.x; i < T3; i += blockDim.x) {
    const int lx = i % T, ly = (i / T) % T, lz = i / (T * T);
    const int gx = imodulo(ox + lx, Nx), gy = imodulo(oy + ly, Ny), gz = imodulo(oz + lz, Nz);
    const size_t gi = ((size_t)gz * Ny + gy) * Nx + gx;
    tile[i] = a.vv[gi];
    tile[T3 + i] = a.vv[a.G + gi];
    tile[2 * T3 + i] = a.vv[2 * a.G + gi];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *w = sw + warp * 3 * PME_PMAX;
  for (int s = sb + warp; s < se; s += INTERP_WARPS) {
    const int p = a.order[s];
    int rel = 0;
    __syncwarp();
    if (lane < 3) {
      const double u = __dmul_rn(a.x[(size_t)lane * a.n + p], a.prm.ih[lane]);
      int imin;
      double ww[PME_PMAX];
      bspline_func<PME_PMAX>(u, P, imin, ww);
      const int mcell = imodulo(imin + (P - 1), a.prm.Nb[lane]);
      const int b = (lane == 0) ? bx : (lane == 1) ? by : bz;
      rel = mcell - b * PME_BLK;
      for (int q = 0; q < P; q++) w[lane * PME_PMAX + q] = ww[q];
    }
    __syncwarp();
    const int rx = __shfl_sync(FULL_MASK, rel, 0), ry = __shfl_sync(FULL_MASK, rel, 1),
              rz = __shfl_sync(FULL_MASK, rel, 2);
    double d0 = 0, d1 = 0, d2 = 0;
    for (int pr = lane; pr < P * P; pr += 32) {
      const int j0 = pr % P, k0 = pr / P;
      const double wyz = w[PME_PMAX + j0] * w[2 * PME_PMAX + k0];
      const int base = ((rz + k0) * T + (ry + j0)) * T + rx;
      for (int i0 = 0; i0 < P; i0++) {
        const double wt = w[i0] * wyz;
        d0 += wt * tile[base + i0];
        d1 += wt * tile[T3 + base + i0];
        d2 += wt * tile[2 * T3 + base + i0];
      }
    }
    d0 = warp_sum(d0);
    d1 = warp_sum(d1);
    d2 = warp_sum(d2);
    if (lane == 0) {
      a.acc[p] += d0;
      a.acc[(size_t)a.n + p] += d1;
      a.acc[2 * (size_t)a.n + p] += d2;
    }
  }
}


// ---------------------------------------------------------------------------------------------------------
// P = 8 interpolation, v2: register-ring column walk.  One warp per (x, y) column of mesh cells; its targets are
// sorted by z cell.  The 8 x 8 x 8 support of a mesh cell lives in REGISTERS: lane = (x offset, y offset pair) holds a
// ring of the last 8 z planes of its two (x, y) columns for the 3 velocity components (48 doubles).  Moving up one
// cell in z replaces one ring slot (64-byte row segments from L1/L2); the ring position is warp-uniform, so a switch
// selects one of 8 bodies with static register indices.  Per target and lane: 48 FMAs against the z weights, 9
// multiply-adds with w_x w_y, then a transposed shuffle reduction over 4 targets at a time -- no shared-memory
// traffic per mesh value at all (v1 needed one LDS per FMA).  The B-spline weights are geometry-time records
// (celllist_pme_weights) staged by wk_stage; results leave as one FP64 reduction per target and component.
constexpr int IW_WARPS = 4;

struct InterpWArgs {
  Params prm;
  int n;
  const int *start, *order;  // mesh-cell list, key = cz + Nz * (cx + Nx * cy)
  const double *wrec;        // [sorted target][PME_WREC]
  const double *vv;          // [3][G]
  size_t G;
  double *acc;               // SoA(3,n)
};

// up to 4 targets against the ring: slot d holds plane cz - 7 + ((d - cz + 7) & 7); the records carry the z weights
// already in ring order, so the register indices are static -- one body for all ring positions
__device__ __forceinline__ void iw_batch(const double (&R)[8][2][3], const double *__restrict__ w, int nb,
                                         int ix, int jq, double (&v)[12]) {
#pragma unroll
  for (int q = 0; q < 4; q++) {
    if (q < nb) {
      const double *wq = w + q * WK_REC;
      double wz[8];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const double2 t = reinterpret_cast<const double2 *>(wq + 16)[k];
        wz[2 * k] = t.x;
        wz[2 * k + 1] = t.y;
      }
      const double wxl = wq[ix], wy0 = wq[8 + jq], wy1 = wq[12 + jq];
#pragma unroll
      for (int cc = 0; cc < 3; cc++) {
        double t0 = 0.0, t1 = 0.0;
#pragma unroll
        for (int d = 0; d < 8; d++) {
          t0 = fma(wz[d], R[d][0][cc], t0);
          t1 = fma(wz[d], R[d][1][cc], t1);
        }
        v[3 * q + cc] = wxl * fma(wy1, t1, wy0 * t0);
      }
    } else {
      v[3 * q] = v[3 * q + 1] = v[3 * q + 2] = 0.0;
    }
  }
}

__global__ void __launch_bounds__(IW_WARPS * 32, 3) k_interp_walk(InterpWArgs a) {
  __shared__ __align__(16) double s_rec[IW_WARPS][2][WK_ROUND * WK_REC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Nx = a.prm.Nb[0], Ny = a.prm.Nb[1], Nz = a.prm.Nb[2];
  const int col = blockIdx.x * IW_WARPS + warp;
  if (col >= Nx * Ny) return;
  const int *st = a.start + (size_t)col * Nz;
  const int pos0 = st[0], end = st[Nz];
  if (pos0 == end) return;
  wk_stage(s_rec[warp][0], lane, pos0, end - 1, a.wrec, nullptr, a.order);
  const int cx = col % Nx, cy = col / Nx;
  const int ix = lane & 7, jq = lane >> 3;
  const int gx = imodulo(cx - 7 + ix, Nx);
  const size_t off0 = (size_t)imodulo(cy - 7 + jq, Ny) * Nx + gx, off1 = (size_t)imodulo(cy - 3 + jq, Ny) * Nx + gx;
  const size_t plane = (size_t)Nx * Ny;
  double R[8][2][3];
  int pos = pos0, cz = 0, sprev = -64, cur_round = -1;
  int zwin = 0;                                          // scan window: e = st[zwin + lane + 1]
  int e = (lane < Nz) ? st[lane + 1] : end + 1;
  while (pos < end) {
    int cnt;
    for (;;) {  // next occupied z cell of the column
      const unsigned m = __ballot_sync(FULL_MASK, e > pos);
      if (m) {
        const int f = __ffs(m) - 1;
        cnt = __shfl_sync(FULL_MASK, e, f) - pos;
        cz = zwin + f;
        break;
      }
      zwin += 32;
      e = (zwin + lane < Nz) ? st[zwin + lane + 1] : end + 1;
    }
    const int b = cz & 7;
#pragma unroll
    for (int d = 0; d < 8; d++) {  // ring slots that are not yet loaded: planes (sprev, cz]
      const int ud = cz - ((b - d) & 7);
      if (ud > sprev) {
        const double *pl = a.vv + (size_t)imodulo(ud, Nz) * plane;
#pragma unroll
        for (int cc = 0; cc < 3; cc++) {
          R[d][0][cc] = __ldg(pl + cc * a.G + off0);
          R[d][1][cc] = __ldg(pl + cc * a.G + off1);
        }
      }
    }
    sprev = cz;
    for (int tb = 0; tb < cnt;) {
      const int rel = pos + tb - pos0, r = rel >> 3, slot = rel & 7;
      if (r != cur_round) {  // warp-uniform: enter round r (staged one round ahead), prefetch round r + 1
        cp_async_wait_all();
        __syncwarp();
        cur_round = r;
        const int s1 = pos0 + 8 * (r + 1);
        if (s1 < end) wk_stage(s_rec[warp][(r + 1) & 1], lane, s1, end - 1, a.wrec, nullptr, a.order);
      }
      const int nb = min(min(4, cnt - tb), WK_ROUND - slot);
      const double *w = s_rec[warp][r & 1] + slot * WK_REC;
      double v[12];
      iw_batch(R, w, nb, ix, jq, v);
      // transposed reduction: 12 -> 6 -> 3 values per lane, then three butterflies; lanes 8q..8q+7 end with target q
      const bool hiA = (lane & 16) != 0, hiB = (lane & 8) != 0;
      double r6[6], r3[3];
#pragma unroll
      for (int k = 0; k < 6; k++) {
        const double send = hiA ? v[k] : v[k + 6], keep = hiA ? v[k + 6] : v[k];
        r6[k] = keep + __shfl_xor_sync(FULL_MASK, send, 16);
      }
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const double send = hiB ? r6[k] : r6[k + 3], keep = hiB ? r6[k + 3] : r6[k];
        r3[k] = keep + __shfl_xor_sync(FULL_MASK, send, 8);
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1)
#pragma unroll
        for (int k = 0; k < 3; k++) r3[k] += __shfl_xor_sync(FULL_MASK, r3[k], o);
      if (ix == 0 && jq < nb) {  // single writer per address and kernel: a reduction without return value
        const int p = reinterpret_cast<const int *>(w + jq * WK_REC + 26)[0];
        atomicAdd(a.acc + p, r3[0]);
        atomicAdd(a.acc + (size_t)a.n + p, r3[1]);
        atomicAdd(a.acc + 2 * (size_t)a.n + p, r3[2]);
      }
      tb += nb;
    }
    pos += cnt;
  }
}

// Short target lists (the wall vertices of examples/minicase, raw point sets): one warp per target straight from the
// mesh, no list walk.  A column walk is a serial chain per (x, y) column whatever the list length; 12 KB of L2 reads
// per target are cheaper than that up to a few thousand targets.  Lane = two (y, z) lines of the 8 x 8 x 8 support.
__global__ void __launch_bounds__(256) k_interp_direct(int n, const double *__restrict__ x, const int *__restrict__ active,
                                                       Params prm, const double *__restrict__ vv, size_t G,
                                                       double *__restrict__ acc) {
  __shared__ double s_w[8][3][8];
  __shared__ int s_i[8][3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = blockIdx.x * 8 + warp;
  if (p >= n || !active[p]) return;
  if (lane < 3) {
    int imin;
    double ww[8];
    bspline_func<8>(__dmul_rn(x[(size_t)lane * n + p], prm.ih[lane]), 8, imin, ww);
#pragma unroll
    for (int q = 0; q < 8; q++) s_w[warp][lane][q] = ww[q];
    s_i[warp][lane] = imin;
  }
  __syncwarp();
  const int Nx = prm.Nb[0], Ny = prm.Nb[1], Nz = prm.Nb[2];
  int gx[8];
#pragma unroll
  for (int q = 0; q < 8; q++) gx[q] = imodulo(s_i[warp][0] + q, Nx);
  double r[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int jy = lane & 7, jz = (lane >> 3) + 4 * h;
    const double *row = vv + ((size_t)imodulo(s_i[warp][2] + jz, Nz) * Ny + imodulo(s_i[warp][1] + jy, Ny)) * Nx;
    const double wyz = s_w[warp][1][jy] * s_w[warp][2][jz];
#pragma unroll
    for (int cc = 0; cc < 3; cc++) {
      double t = 0.0;
#pragma unroll
      for (int q = 0; q < 8; q++) t = fma(s_w[warp][0][q], __ldg(row + cc * G + gx[q]), t);
      r[cc] = fma(wyz, t, r[cc]);
    }
  }
#pragma unroll
  for (int cc = 0; cc < 3; cc++) r[cc] = warp_sum(r[cc]);
  if (lane == 0) {
    acc[p] += r[0];
    acc[(size_t)n + p] += r[1];
    acc[2 * (size_t)n + p] += r[2];
  }
}

int pme_interp(rbc3d_ctx *c, TargetList &t, double *acc) {
  Pme &pm = c->pme;
  if (!pm.transformed) {
    set_error("PME_Add_Interp_Vel called before PME_Transform");
    return RBC3D_ESTATE;
  }
  if (pm.slab && c->prm.nranks > 1) RBC_TRY(pme_halo_exchange(c, t));  // collective: before the empty-list return
  if (t.n == 0) return RBC3D_OK;
  CellList &pl = t.pl;
  if (pm.walk && c->prm.nranks == 1 && t.n <= pm.interp_direct_max) {
    k_interp_direct<<<(t.n + 7) / 8, 256, 0, c->stream>>>(t.n, t.x.p, t.active.p, c->prm, pm.vv.p, pm.G, acc ? acc : t.acc.p);
    KERNEL_CHECK();
    c->launches++;
    return RBC3D_OK;
  }
  if (pm.walk) {
    InterpWArgs w;
    w.prm = c->prm;
    w.n = t.n;
    w.start = pl.start.p;
    w.order = pl.order.p;
    w.wrec = pl.w.p;
    w.vv = pm.vv.p;
    w.G = pm.G;
    w.acc = acc ? acc : t.acc.p;
    const int ncol = pm.Nx * pm.Ny;
    k_interp_walk<<<(ncol + IW_WARPS - 1) / IW_WARPS, IW_WARPS * 32, 0, c->stream>>>(w);
    KERNEL_CHECK();
    c->launches++;
    return RBC3D_OK;
  }
  InterpArgs a;
  a.prm = c->prm;
  a.n = t.n;
  a.start = pl.start.p;
  a.order = pl.order.p;
  a.x = t.x.p;
  a.vv = pm.vv.p;
  a.G = pm.G;
  a.nbx = pm.nblk[0];
  a.nby = pm.nblk[1];
  a.nbz = pm.nblk[2];
  a.acc = acc ? acc : t.acc.p;
  const int T = PME_BLK + c->prm.P - 1;
  const size_t smem = sizeof(double) * (3 * (size_t)T * T * T + INTERP_WARPS * 3 * PME_PMAX);
  CUDA_TRY(cudaFuncSetAttribute(k_interp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_interp<<<a.nbx * a.nby * a.nbz, INTERP_WARPS * 32, smem, c->stream>>>(a);
  KERNEL_CHECK();
  c->launches++;
  return RBC3D_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Slab-decomposed PME (several ranks): ModPFFTW.F90:56-89 (Init_PFFTW), :94-185 (transforms), :188-316 (transposes);
// ModPME.F90:354-396 (Update_Buff_Vel), :428-429 (spreading skips planes of other ranks); ModConf.F90:412-437.
//
//   z-slab of rank r: planes [zoff[r], zoff[r+1]);  y-slab: rows [yoff[r], yoff[r+1]) of the half spectrum.
//   forward :  2-D D2Z per owned plane (cuFFT, batch = planes)  ->  pack per destination [comp][z of sender][y of
//              receiver][Nxh]  ->  grouped ncclSend/ncclRecv  ->  unpack + transpose to [comp][y][Nxh][Nz] (z fastest)
//              ->  1-D Z2Z in z (cuFFT, contiguous)  ->  k-space multiplier on the y-slab
//   backward:  1-D Z2Z inverse  ->  pack + transpose  ->  exchange  ->  unpack into the owned planes of the full-size
//              spectral array  ->  Hermitian symmetrisation of the columns x = 0, Nx/2 in k_y, plane by plane (what FFTW's
//              c2r does implicitly and the single-rank path does in (k_y, k_z) before its 3-D Z2D: no partner rank is
//              needed)  ->  2-D Z2D per owned plane.
// Real and spectral meshes keep their full-size single-rank layout [comp][Nz][Ny][..]; a rank only touches its planes
// (+ the halo planes of the velocity mesh), so spreading and interpolation kernels are the single-rank ones.
static int slab_forced() {
  static const int v = [] {
    const char *e = getenv("RBC3D_PME_SLAB");  // 1: run the decomposed transform on one rank too (tests)
    return e ? atoi(e) : 0;
  }();
  return v;
}

int pme_slab_setup(rbc3d_ctx *c) {
  Pme &pm = c->pme;
  const int R = c->prm.nranks;
  static const bool off = getenv("RBC3D_PME_ALLREDUCE") != nullptr;  // round-1 behaviour: full meshes summed with one all-reduce
  pm.slab = (R > 1 && !off) || slab_forced() == 1;
  pm.R = R, pm.rk = c->prm.rank;
  pm.zoff.assign(R + 1, 0), pm.yoff.assign(R + 1, 0);
  for (int r = 0; r <= R; r++) {
    pm.zoff[r] = (int)((long long)pm.Nz * r / R);
    pm.yoff[r] = (int)((long long)pm.Ny * r / R);
  }
  if (!pm.slab) return RBC3D_OK;
  RBC_TRY(pm.d_zoff.resize(R + 1));
  RBC_TRY(pm.d_yoff.resize(R + 1));
  CUDA_TRY(cudaMemcpy(pm.d_zoff.p, pm.zoff.data(), sizeof(int) * (R + 1), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(pm.d_yoff.p, pm.yoff.data(), sizeof(int) * (R + 1), cudaMemcpyHostToDevice));
  return RBC3D_OK;
}

// sources a rank spreads: every point whose B-spline support (planes cz - P + 1 .. cz) touches the rank's planes
__global__ void k_slab_own(int n, const double *__restrict__ z, double ihz, int Nz, int z0, int nz, int P, int *__restrict__ own) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cz = imodulo((int)floor(__dmul_rn(z[i], ihz)), Nz);  // mesh cell of the point, as the spreading kernels
  own[i] = imodulo(cz - z0, Nz) < nz + P - 1 ? 1 : 0;
}

int pme_source_ownership(rbc3d_ctx *c, int n, const double *x, dbuf<int> &own, const int **flags) {
  Pme &pm = c->pme;
  *flags = nullptr;
  if (c->prm.nranks <= 1 || n == 0) return RBC3D_OK;
  RBC_TRY(own.resize(n));
  if (pm.slab) {
    const int z0 = pm.zoff[pm.rk], nz = pm.zoff[pm.rk + 1] - z0;
    k_slab_own<<<(n + 255) / 256, 256, 0, c->stream>>>(n, x + 2 * (size_t)n, c->prm.ih[2], pm.Nz, z0, nz, c->prm.P, own.p);
    KERNEL_CHECK();
    c->launches++;
    *flags = own.p;
  }
  return RBC3D_OK;
}

struct SlabArgs {
  int R, rk, nc, Nz, Ny, Nxh;
  const int *zoff, *yoff;
  size_t M;
};

__device__ __forceinline__ int slab_owner(const int *off, int R, int v) {  // rank r with off[r] <= v < off[r+1]
  int r = 0;
  while (r + 1 < R && v >= off[r + 1]) r++;
  return r;
}

// element offset of the block exchanged between (z-slab owner zr, y-slab owner yr) inside the buffer of one of them:
// blocks are ordered by the PEER rank; a block holds [comp][planes of zr][rows of yr][Nxh]
__device__ __forceinline__ size_t slab_block_off(const SlabArgs &a, int zr, int yr, bool at_z_owner) {
  size_t o = 0;
  if (at_z_owner) {  // buffer of the z owner: one block per y owner
    const size_t nz = a.zoff[zr + 1] - a.zoff[zr];
    o = (size_t)a.nc * nz * a.yoff[yr] * a.Nxh;
  } else {           // buffer of the y owner: one block per z owner
    const size_t ny = a.yoff[yr + 1] - a.yoff[yr];
    o = (size_t)a.nc * a.zoff[zr] * ny * a.Nxh;
  }
  return o;
}

// forward pack: the rank's planes of the 2-D spectra [comp][Nz][Ny][Nxh] -> send blocks (one thread per element, x fastest)
__global__ void __launch_bounds__(256) k_slab_pack_fwd(SlabArgs a, const cufftDoubleComplex *__restrict__ src,
                                                       cufftDoubleComplex *__restrict__ buf) {
  const int z0 = a.zoff[a.rk], nz = a.zoff[a.rk + 1] - z0;
  const size_t total = (size_t)a.nc * nz * a.Ny * a.Nxh;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(e % a.Nxh);
    size_t r = e / a.Nxh;
    const int y = (int)(r % a.Ny);
    r /= a.Ny;
    const int zl = (int)(r % nz), comp = (int)(r / nz);
    const int s = slab_owner(a.yoff, a.R, y);
    const int nys = a.yoff[s + 1] - a.yoff[s];
    const size_t d = slab_block_off(a, a.rk, s, true) + (((size_t)comp * nz + zl) * nys + (y - a.yoff[s])) * a.Nxh + x;
    buf[d] = src[(size_t)comp * a.M + ((size_t)(z0 + zl) * a.Ny + y) * a.Nxh + x];
  }
}

// forward unpack: received blocks [z owner][comp][its planes][my rows][Nxh] -> Tz[comp][my rows][Nxh][Nz] (z fastest);
// 32 x 32 tiles over (z, x) through shared memory, grid (x tiles, z tiles, comp * my rows)
__global__ void __launch_bounds__(256) k_slab_unpack_fwd(SlabArgs a, const cufftDoubleComplex *__restrict__ buf,
                                                         cufftDoubleComplex *__restrict__ Tz) {
  __shared__ cufftDoubleComplex tile[32][33];
  const int ny = a.yoff[a.rk + 1] - a.yoff[a.rk];
  const int comp = blockIdx.z / ny, yl = blockIdx.z - comp * ny;
  const int x0 = blockIdx.x * 32, zt0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int z = zt0 + r, x = x0 + threadIdx.x;
    if (z < a.Nz && x < a.Nxh) {
      const int s = slab_owner(a.zoff, a.R, z);
      const int nzs = a.zoff[s + 1] - a.zoff[s];
      tile[r][threadIdx.x] =
          buf[slab_block_off(a, s, a.rk, false) + (((size_t)comp * nzs + (z - a.zoff[s])) * ny + yl) * a.Nxh + x];
    }
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int x = x0 + r, z = zt0 + threadIdx.x;
    if (z < a.Nz && x < a.Nxh) Tz[(((size_t)comp * ny + yl) * a.Nxh + x) * a.Nz + z] = tile[threadIdx.x][r];
  }
}

// backward pack: Vz[comp][my rows][Nxh][Nz] -> send blocks [z owner][comp][its planes][my rows][Nxh]
__global__ void __launch_bounds__(256) k_slab_pack_bwd(SlabArgs a, const cufftDoubleComplex *__restrict__ Vz,
                                                       cufftDoubleComplex *__restrict__ buf) {
  __shared__ cufftDoubleComplex tile[32][33];
  const int ny = a.yoff[a.rk + 1] - a.yoff[a.rk];
  const int comp = blockIdx.z / ny, yl = blockIdx.z - comp * ny;
  const int x0 = blockIdx.x * 32, zt0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int x = x0 + r, z = zt0 + threadIdx.x;
    if (z < a.Nz && x < a.Nxh) tile[r][threadIdx.x] = Vz[(((size_t)comp * ny + yl) * a.Nxh + x) * a.Nz + z];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int z = zt0 + r, x = x0 + threadIdx.x;
    if (z < a.Nz && x < a.Nxh) {
      const int s = slab_owner(a.zoff, a.R, z);
      const int nzs = a.zoff[s + 1] - a.zoff[s];
      buf[slab_block_off(a, s, a.rk, false) + (((size_t)comp * nzs + (z - a.zoff[s])) * ny + yl) * a.Nxh + x] =
          tile[threadIdx.x][r];
    }
  }
}

// backward unpack: received blocks [y owner][comp][my planes][its rows][Nxh] -> my planes of [comp][Nz][Ny][Nxh]
__global__ void __launch_bounds__(256) k_slab_unpack_bwd(SlabArgs a, const cufftDoubleComplex *__restrict__ buf,
                                                         cufftDoubleComplex *__restrict__ dst) {
  const int z0 = a.zoff[a.rk], nz = a.zoff[a.rk + 1] - z0;
  const size_t total = (size_t)a.nc * nz * a.Ny * a.Nxh;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(e % a.Nxh);
    size_t r = e / a.Nxh;
    const int y = (int)(r % a.Ny);
    r /= a.Ny;
    const int zl = (int)(r % nz), comp = (int)(r / nz);
    const int s = slab_owner(a.yoff, a.R, y);
    const int nys = a.yoff[s + 1] - a.yoff[s];
    dst[(size_t)comp * a.M + ((size_t)(z0 + zl) * a.Ny + y) * a.Nxh + x] =
        buf[slab_block_off(a, a.rk, s, true) + (((size_t)comp * nz + zl) * nys + (y - a.yoff[s])) * a.Nxh + x];
  }
}

// FFTW's c2r over (x, y) uses only the real part of the x = 0 and x = Nx/2 bins after the y transform
// (ModPFFTW.F90:139-143); with cuFFT's Z2D the same result needs those two columns Hermitian in k_y: V(j) <- (V(j) +
// conj V(Ny - j)) / 2, plane by plane.  One thread per (comp, plane, column, j <= Ny/2).
__global__ void k_slab_sym(SlabArgs a, int Nx, cufftDoubleComplex *__restrict__ v) {
  const int z0 = a.zoff[a.rk], nz = a.zoff[a.rk + 1] - z0, nh = a.Ny / 2 + 1;
  const size_t total = (size_t)a.nc * nz * 2 * nh, e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int j = (int)(e % nh);
  size_t r = e / nh;
  const int col = (int)(r % 2);
  r /= 2;
  const int zl = (int)(r % nz), comp = (int)(r / nz);
  const int x = col == 0 ? 0 : Nx / 2;
  if (col == 1 && (Nx & 1)) return;  // odd Nx has no Nyquist column
  cufftDoubleComplex *pl = v + (size_t)comp * a.M + (size_t)(z0 + zl) * a.Ny * a.Nxh + x;
  const int jm = (a.Ny - j) % a.Ny;
  if (jm < j) return;  // the pair is handled by its smaller index
  const cufftDoubleComplex p = pl[(size_t)j * a.Nxh], q = pl[(size_t)jm * a.Nxh];
  const double re = 0.5 * (p.x + q.x), im = 0.5 * (p.y - q.y);
  pl[(size_t)j * a.Nxh] = make_cuDoubleComplex(re, im);
  pl[(size_t)jm * a.Nxh] = make_cuDoubleComplex(re, jm == j ? 0.0 : -im);
}

template <bool SL, bool DL>
__global__ void __launch_bounds__(256) k_kspace_slab(KArgs a, int ny) {  // one thread per mode of the y-slab, z fastest
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)ny * a.Nxh * a.Nz;
  if (idx >= total) return;
  const int k = (int)(idx % a.Nz);
  const int i = (int)((idx / a.Nz) % a.Nxh);
  const int j = a.j0 + (int)(idx / ((size_t)a.Nz * a.Nxh));
  double vr[3], vi[3];
  kspace_mode<SL, DL>(a, i, j, k, vr, vi);
#pragma unroll
  for (int d = 0; d < 3; d++) a.vvC[(size_t)d * total + idx] = make_cuDoubleComplex(vr[d], vi[d]);
}

// all-to-all of blocks between z owners and y owners.  to_y: every rank sends block (me, s) of sbuf to s and receives
// block (s, me) into rbuf; otherwise the reverse direction.  Block sizes differ when Nz or Ny is not a multiple of R.
static int slab_exchange(rbc3d_ctx *c, int nc, bool to_y) {
  Pme &pm = c->pme;
  const int R = pm.R, me = pm.rk;
  const size_t nzm = pm.zoff[me + 1] - pm.zoff[me], nym = pm.yoff[me + 1] - pm.yoff[me];
  auto blk = [&](int zr, int yr) {
    return (size_t)nc * (pm.zoff[zr + 1] - pm.zoff[zr]) * (pm.yoff[yr + 1] - pm.yoff[yr]) * pm.Nxh;
  };
  if (R > 1) RBC_TRY(comm_group_begin());
  for (int s = 0; s < R; s++) {
    // to_y: I am the z owner on the send side (blocks ordered by y owner), the y owner on the receive side
    const size_t soff = to_y ? (size_t)nc * nzm * pm.yoff[s] * pm.Nxh : (size_t)nc * pm.zoff[s] * nym * pm.Nxh;
    const size_t roff = to_y ? (size_t)nc * pm.zoff[s] * nym * pm.Nxh : (size_t)nc * nzm * pm.yoff[s] * pm.Nxh;
    const size_t ns = to_y ? blk(me, s) : blk(s, me), nr = to_y ? blk(s, me) : blk(me, s);
    if (s == me || R == 1) {
      CUDA_TRY(cudaMemcpyAsync(pm.rbuf.p + roff, pm.sbuf.p + soff, ns * sizeof(cufftDoubleComplex), cudaMemcpyDeviceToDevice,
                               c->stream));
    } else {
      RBC_TRY(comm_send(c, pm.sbuf.p + soff, ns * sizeof(cufftDoubleComplex), s));
      RBC_TRY(comm_recv(c, pm.rbuf.p + roff, nr * sizeof(cufftDoubleComplex), s));
    }
  }
  if (R > 1) RBC_TRY(comm_group_end());
  return RBC3D_OK;
}

static int pme_transform_slab(rbc3d_ctx *c) {
  Pme &pm = c->pme;
  const Params &p = c->prm;
  const int R = pm.R, me = pm.rk;
  const int z0 = pm.zoff[me], nz = pm.zoff[me + 1] - z0, ny = pm.yoff[me + 1] - pm.yoff[me];
  const int nc = (pm.flag_sl ? 3 : 0) + (pm.flag_dl ? 6 : 0), comp0 = pm.flag_sl ? 0 : 3;
  const size_t plane_r = (size_t)pm.Ny * pm.Nx, plane_c = (size_t)pm.Ny * pm.Nxh;
  RBC_TRY(pm.sbuf.resize((size_t)9 * nz * plane_c));
  RBC_TRY(pm.rbuf.resize((size_t)9 * pm.Nz * ny * pm.Nxh));
  RBC_TRY(pm.Tz.resize((size_t)9 * pm.Nz * ny * pm.Nxh));
  RBC_TRY(pm.Vz.resize((size_t)3 * pm.Nz * ny * pm.Nxh));
  if (!pm.plan2F_ok) {
    int n2[2] = {pm.Ny, pm.Nx};
    CUFFT_TRY(cufftPlanMany(&pm.plan2F, 2, n2, nullptr, 1, (int)plane_r, nullptr, 1, (int)plane_c, CUFFT_D2Z, nz));
    CUFFT_TRY(cufftPlanMany(&pm.plan2B, 2, n2, nullptr, 1, (int)plane_c, nullptr, 1, (int)plane_r, CUFFT_Z2D, nz));
    pm.plan2F_ok = pm.plan2B_ok = true;
  }
  auto plan1 = [&](int slot, int batch, cufftHandle *out) -> int {
    if (pm.plan1_ok[slot] && pm.plan1_batch[slot] != batch) {
      cufftDestroy(pm.plan1[slot]);
      pm.plan1_ok[slot] = false;
    }
    if (!pm.plan1_ok[slot]) {
      int n1[1] = {pm.Nz};
      CUFFT_TRY(cufftPlanMany(&pm.plan1[slot], 1, n1, nullptr, 1, pm.Nz, nullptr, 1, pm.Nz, CUFFT_Z2Z, batch));
      pm.plan1_ok[slot] = true;
      pm.plan1_batch[slot] = batch;
    }
    *out = pm.plan1[slot];
    CUFFT_TRY(cufftSetStream(pm.plan1[slot], c->stream));
    return RBC3D_OK;
  };
  CUFFT_TRY(cufftSetStream(pm.plan2F, c->stream));
  CUFFT_TRY(cufftSetStream(pm.plan2B, c->stream));
  SlabArgs sa;
  sa.R = R, sa.rk = me, sa.nc = nc, sa.Nz = pm.Nz, sa.Ny = pm.Ny, sa.Nxh = pm.Nxh;
  sa.zoff = pm.d_zoff.p, sa.yoff = pm.d_yoff.p, sa.M = pm.M;
  // ---- forward ----
  t_begin(c, RBC3D_T_FFT);
  for (int q = 0; q < nc; q++)
    CUFFT_TRY(cufftExecD2Z(pm.plan2F, pm.src.p + (size_t)(comp0 + q) * pm.G + (size_t)z0 * plane_r,
                           pm.srcC.p + (size_t)(comp0 + q) * pm.M + (size_t)z0 * plane_c));
  k_slab_pack_fwd<<<c->sm_count * 8, 256, 0, c->stream>>>(sa, pm.srcC.p + (size_t)comp0 * pm.M, pm.sbuf.p);
  KERNEL_CHECK();
  t_end(c, RBC3D_T_FFT);
  t_begin(c, RBC3D_T_COMM);
  RBC_TRY(slab_exchange(c, nc, true));
  t_end(c, RBC3D_T_COMM);
  t_begin(c, RBC3D_T_KSPACE);
  const dim3 tb(32, 8), tg((pm.Nxh + 31) / 32, (pm.Nz + 31) / 32, nc * ny);
  if (ny > 0) {
    k_slab_unpack_fwd<<<tg, tb, 0, c->stream>>>(sa, pm.rbuf.p, pm.Tz.p);
    KERNEL_CHECK();
    cufftHandle p1;
    RBC_TRY(plan1(0, nc * ny * pm.Nxh, &p1));
    CUFFT_TRY(cufftExecZ2Z(p1, pm.Tz.p, pm.Tz.p, CUFFT_FORWARD));
    KArgs a;
    a.prm = p;
    a.Nx = pm.Nx, a.Ny = pm.Ny, a.Nz = pm.Nz, a.Nxh = pm.Nxh;
    a.M = (size_t)ny * pm.Nxh * pm.Nz;                        // component stride on the y-slab
    a.srcC = pm.Tz.p - (size_t)comp0 * a.M;                   // kspace_mode addresses components 0..2 (SL), 3..8 (DL)
    a.vvC = pm.Vz.p;
    a.bx = pm.bx.p, a.by = pm.by.p, a.bz = pm.bz.p;
    a.vol = p.Lb[0] * p.Lb[1] * p.Lb[2];
    a.sx = pm.Nz, a.sy = (long long)pm.Nxh * pm.Nz, a.sz = 1, a.j0 = pm.yoff[me];
    const int grid = (int)((a.M + 255) / 256);
    if (pm.flag_sl && pm.flag_dl)
      k_kspace_slab<true, true><<<grid, 256, 0, c->stream>>>(a, ny);
    else if (pm.flag_sl)
      k_kspace_slab<true, false><<<grid, 256, 0, c->stream>>>(a, ny);
    else
      k_kspace_slab<false, true><<<grid, 256, 0, c->stream>>>(a, ny);
    KERNEL_CHECK();
    // ---- backward ----
    RBC_TRY(plan1(1, 3 * ny * pm.Nxh, &p1));
    CUFFT_TRY(cufftExecZ2Z(p1, pm.Vz.p, pm.Vz.p, CUFFT_INVERSE));
    sa.nc = 3;
    const dim3 tg3((pm.Nxh + 31) / 32, (pm.Nz + 31) / 32, 3 * ny);
    k_slab_pack_bwd<<<tg3, tb, 0, c->stream>>>(sa, pm.Vz.p, pm.sbuf.p);
    KERNEL_CHECK();
  }
  sa.nc = 3;
  t_end(c, RBC3D_T_KSPACE);
  RBC_TRY(slab_exchange(c, 3, false));
  t_begin(c, RBC3D_T_FFT_INV);
  k_slab_unpack_bwd<<<c->sm_count * 8, 256, 0, c->stream>>>(sa, pm.rbuf.p, pm.vvC.p);
  {
    const size_t tot = (size_t)3 * nz * 2 * (pm.Ny / 2 + 1);
    if (tot) k_slab_sym<<<(int)((tot + 255) / 256), 256, 0, c->stream>>>(sa, pm.Nx, pm.vvC.p);
  }
  KERNEL_CHECK();
  for (int d = 0; d < 3; d++)
    CUFFT_TRY(cufftExecZ2D(pm.plan2B, pm.vvC.p + (size_t)d * pm.M + (size_t)z0 * plane_c,
                           pm.vv.p + (size_t)d * pm.G + (size_t)z0 * plane_r));
  t_end(c, RBC3D_T_FFT_INV);
  c->launches += 8 + nc + 3;
  pm.transformed = true;
  return RBC3D_OK;
}

// ---- velocity-mesh halo (Update_Buff_Vel, ModPME.F90:354-396): the planes a rank's active targets interpolate from ----
// per active target the signed plane offset of its mesh cell from the first owned plane; min / max over the list
__global__ void k_halo_range(int n, const double *__restrict__ z, const int *__restrict__ active, double ihz, int Nz, int z0,
                             int *__restrict__ mm) {
  int lo = 1 << 30, hi = -(1 << 30);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (active && !active[i]) continue;
    const int cz = imodulo((int)floor(__dmul_rn(z[i], ihz)), Nz);
    const int d = imodulo(cz - z0 + Nz / 2, Nz) - Nz / 2;
    lo = min(lo, d), hi = max(hi, d);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(FULL_MASK, lo, o));
    hi = max(hi, __shfl_xor_sync(FULL_MASK, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(mm, lo);
    atomicMax(mm + 1, hi);
  }
}

static int pme_halo_exchange(rbc3d_ctx *c, TargetList &t) {
  Pme &pm = c->pme;
  const int R = pm.R, me = pm.rk, P = c->prm.P;
  const int z0 = pm.zoff[me];
  if (t.halo_version != t.version || (int)t.halo_need.size() != 2 * R) {
    // needed planes of this list on this rank: [z0 + dmin - (P - 1), z0 + dmax], then the same of every rank
    RBC_TRY(pm.halo_tmp.resize(2 + 2 * (size_t)R));
    int init[2] = {1 << 30, -(1 << 30)};
    CUDA_TRY(cudaMemcpyAsync(pm.halo_tmp.p, init, sizeof init, cudaMemcpyHostToDevice, c->stream));
    if (t.n > 0) {
      k_halo_range<<<std::min(c->sm_count * 4, (t.n + 255) / 256), 256, 0, c->stream>>>(t.n, t.x.p + 2 * (size_t)t.n, t.active.p,
                                                                                         c->prm.ih[2], pm.Nz, z0, pm.halo_tmp.p);
      KERNEL_CHECK();
    }
    int mm[2];
    CUDA_TRY(cudaMemcpyAsync(mm, pm.halo_tmp.p, sizeof mm, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    int need[2] = {0, 0};  // (first plane, count); count 0 = no active target
    if (mm[0] <= mm[1]) {
      const int len = std::min(pm.Nz, mm[1] - mm[0] + P);
      need[0] = ((z0 + mm[0] - (P - 1)) % pm.Nz + pm.Nz) % pm.Nz;
      need[1] = len;
    }
    CUDA_TRY(cudaMemcpyAsync(pm.halo_tmp.p, need, sizeof need, cudaMemcpyHostToDevice, c->stream));
    RBC_TRY(comm_allgather_ints(c, pm.halo_tmp.p, pm.halo_tmp.p + 2, 2));
    t.halo_need.assign(2 * (size_t)R, 0);
    CUDA_TRY(cudaMemcpyAsync(t.halo_need.data(), pm.halo_tmp.p + 2, sizeof(int) * 2 * R, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    t.halo_version = t.version;
  }
  // planes of owner `o` inside the needed range of rank `r`: up to two runs of the cyclic range
  const size_t plane = (size_t)pm.Ny * pm.Nx;
  auto runs = [&](int r, int o, int (&seg)[2][2]) {
    int nseg = 0;
    const int lo = t.halo_need[2 * r], len = t.halo_need[2 * r + 1];
    if (len <= 0) return 0;
    const int part[2][2] = {{lo, std::min(lo + len, pm.Nz)}, {0, std::max(0, lo + len - pm.Nz)}};
    for (int k = 0; k < 2; k++) {
      const int a = std::max(part[k][0], pm.zoff[o]), b = std::min(part[k][1], pm.zoff[o + 1]);
      if (a < b) seg[nseg][0] = a, seg[nseg][1] = b, nseg++;
    }
    return nseg;
  };
  t_begin(c, RBC3D_T_COMM);
  RBC_TRY(comm_group_begin());
  for (int s = 0; s < R; s++) {
    if (s == me) continue;
    int seg[2][2];
    int ns = runs(s, me, seg);  // my planes that rank s needs
    for (int k = 0; k < ns; k++)
      for (int d = 0; d < 3; d++)
        RBC_TRY(comm_send(c, pm.vv.p + (size_t)d * pm.G + (size_t)seg[k][0] * plane, (size_t)(seg[k][1] - seg[k][0]) * plane * 8, s));
    ns = runs(me, s, seg);      // planes of rank s that I need
    for (int k = 0; k < ns; k++)
      for (int d = 0; d < 3; d++)
        RBC_TRY(comm_recv(c, pm.vv.p + (size_t)d * pm.G + (size_t)seg[k][0] * plane, (size_t)(seg[k][1] - seg[k][0]) * plane * 8, s));
  }
  RBC_TRY(comm_group_end());
  t_end(c, RBC3D_T_COMM);
  return RBC3D_OK;
}

}  // namespace rbc3d
