// splinebuild.cu -- density splines built on the device.
//
// Every GMRES matvec of the reference rebuilds spln_GdetJ of every cell from the new density before it calls the
// operator (MyMatMult, ModVelSolver.F90:560-565 -> Rbc_BuildSurfaceSource(gFlag), ModRbc.F90:785-802):
//   GdetJ = g * detJ on the Gauss mesh
//   ShAnalGau + ShFilter(nlat0) + ShSynthEqu(nlat+1)      spherical-harmonic projection onto degree < nlat0,
//                                                         evaluated on the equally spaced colatitudes i*pi/nlat
//   Spline_Build_on_Sphere (ModSpline.F90:121-142)        doubled-sphere periodic extension, FFT_Diff in theta and phi
// With host-built splines that is 2.16 GB of host->device traffic per matvec at 4096 cells.  The whole chain is
// linear and separable, so it is applied here as small dense operators per cell and variable:
//   phi analysis     F_m(i)  = sum_j gd(i,j) e^{-i m phi_j},  m < nlat0             (DFT rows, shared memory)
//   theta operators  U_m(I)  = sum_i M0[m](I,i) F_m(i),  U1_m = M1[m] F_m           (I = 0..2nlat-1)
//                    M0[m] = E_m T_m with T_m(o,i) = sum_{n<nlat0} Pbar_n^m(cos o pi/nlat) Pbar_n^m(cos th_i) wg_i and
//                    E_m the doubled-sphere extension u(nlat+i, j) = v(nlat-i, j+nlon/2)  (a factor (-1)^m);
//                    M1[m] = D_theta M0[m], D_theta = the circulant of FFT_Diff (ModFFT.F90:25-93, Nyquist zeroed)
//   phi synthesis    u, u1 from U, U1;  u2, u12 by multiplying the modes with i m
// and the result is written both in the ABI layout (direct singular kernel, near-singular kernels) and in the
// double2-plane layout of the cached singular kernel.  Host side: the Legendre tables (once per mesh).
#include <cmath>
#include <vector>

#include "device_math.cuh"
#include "rbc3d_internal.h"

namespace rbc3d {

// fully normalised associated Legendre functions Pbar_n^m(x), 0 <= m <= n < nmax, int_{-1}^{1} Pbar^2 dx = 1
static void pbar_table(int nmax, double x, std::vector<double> &out /* [m*nmax + n] */) {
  out.assign((size_t)nmax * nmax, 0.0);
  const double s = sqrt(fmax(0.0, 1.0 - x * x));
  double pmm = sqrt(0.5);
  for (int m = 0; m < nmax; m++) {
    if (m > 0) pmm = pmm * sqrt((2.0 * m + 1.0) / (2.0 * m)) * s;
    out[(size_t)m * nmax + m] = pmm;
    if (m + 1 < nmax) {
      double p0 = pmm, p1 = sqrt(2.0 * m + 3.0) * x * pmm;
      out[(size_t)m * nmax + m + 1] = p1;
      for (int n = m + 2; n < nmax; n++) {
        const double a = sqrt((4.0 * n * n - 1.0) / ((double)n * n - (double)m * m));
        const double b = sqrt((((double)n - 1.0) * ((double)n - 1.0) - (double)m * m) / (4.0 * ((double)n - 1.0) * ((double)n - 1.0) - 1.0));
        const double p2 = a * (x * p1 - b * p0);
        out[(size_t)m * nmax + n] = p2;
        p0 = p1;
        p1 = p2;
      }
    }
  }
}

int spline_builder_prepare(rbc3d_ctx *c, int nlat0) {
  Cells &C = c->cells;
  C.sb_ok = false;
  if (!C.mesh_set || nlat0 < 1 || nlat0 > C.nlat || nlat0 > C.nlon / 2) return RBC3D_EINVAL;
  const int nlat = C.nlat, nlon = C.nlon, M = 2 * nlat;
  const double PI = 3.14159265358979323846;
  std::vector<double> wg(nlat);
  for (int i = 0; i < nlat; i++) wg[i] = C.h_w[i] / (2.0 * PI / nlon);  // w = Gauss weight * 2pi/nlon, ModRbc.F90:95
  // T[m][o][i]
  std::vector<std::vector<double>> pg(nlat), po(nlat + 1);
  for (int i = 0; i < nlat; i++) pbar_table(nlat0, cos(C.h_th[i]), pg[i]);
  for (int o = 0; o <= nlat; o++) pbar_table(nlat0, cos(o * PI / nlat), po[o]);
  std::vector<double> T((size_t)nlat0 * (nlat + 1) * nlat, 0.0);
  for (int m = 0; m < nlat0; m++)
    for (int o = 0; o <= nlat; o++)
      for (int i = 0; i < nlat; i++) {
        double s = 0;
        for (int n = m; n < nlat0; n++) s += po[o][(size_t)m * nlat0 + n] * pg[i][(size_t)m * nlat0 + n];
        T[((size_t)m * (nlat + 1) + o) * nlat + i] = s * wg[i];
      }
  // theta derivative circulant of FFT_Diff on M points: d[k] = -(2/M) sum_{q=1}^{M/2-1} q sin(2 pi q k / M)
  std::vector<double> d(M, 0.0);
  for (int k = 0; k < M; k++) {
    double s = 0;
    for (int q = 1; q < M / 2; q++) s += q * sin(2.0 * PI * q * k / M);
    d[k] = -(2.0 / M) * s;
  }
  // M0t[m][i][I], M1t[m][i][I]  (I fastest: coalesced for thread = I)
  std::vector<double> M0((size_t)nlat0 * nlat * M), M1((size_t)nlat0 * nlat * M);
  std::vector<double> col(M);
  for (int m = 0; m < nlat0; m++)
    for (int i = 0; i < nlat; i++) {
      const double sgn = (m & 1) ? -1.0 : 1.0;
      for (int I = 0; I < M; I++)
        col[I] = I < nlat ? T[((size_t)m * (nlat + 1) + I) * nlat + i] : sgn * T[((size_t)m * (nlat + 1) + (M - I)) * nlat + i];
      for (int I = 0; I < M; I++) {
        double s = 0;
        for (int J = 0; J < M; J++) s += d[((I - J) % M + M) % M] * col[J];
        M0[((size_t)m * nlat + i) * M + I] = col[I];
        M1[((size_t)m * nlat + i) * M + I] = s;
      }
    }
  // phi tables cos(m phi_j), sin(m phi_j)
  std::vector<double> cs((size_t)2 * nlat0 * nlon);
  for (int m = 0; m < nlat0; m++)
    for (int j = 0; j < nlon; j++) {
      const double ang = 2.0 * PI * (double)((m * j) % nlon) / nlon;
      cs[((size_t)m * nlon + j) * 2] = cos(ang);
      cs[((size_t)m * nlon + j) * 2 + 1] = sin(ang);
    }
  RBC_TRY(C.sb_M0.resize(M0.size()));
  RBC_TRY(C.sb_M1.resize(M1.size()));
  RBC_TRY(C.sb_cs.resize(cs.size()));
  CUDA_TRY(cudaMemcpyAsync(C.sb_M0.p, M0.data(), sizeof(double) * M0.size(), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(C.sb_M1.p, M1.data(), sizeof(double) * M1.size(), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(C.sb_cs.p, cs.data(), sizeof(double) * cs.size(), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  C.sb_nlat0 = nlat0;
  C.sb_ok = true;
  return RBC3D_OK;
}

struct BuildArgs {
  int ncell, npc, nlat, nlon, nlat0, Np;
  const double *dens;   // SoA(3,Np): slist density (already * detJ * w)
  const double *w;      // [nlat]
  const double *M0, *M1, *cs;
  double *sp_abi;       // [cell][4][3][nlon][2 nlat] or null
  double2 *sp_planes;   // [cell][6][2 nlat][nlon] or null
  const int *need;      // [cell] or null: build only flagged cells
  int nvar;             // variables per cell of this field (3, or 1 for detJ)
  int raw;              // 1: dens holds the mesh field itself (x, a3, detJ) -- no division by the Gauss weight
};

// one CTA per (cell, variable)
__global__ void __launch_bounds__(256) k_spline_build(BuildArgs a) {
  extern __shared__ double sm[];
  const int nlat = a.nlat, nlon = a.nlon, M = 2 * nlat, m0 = a.nlat0;
  const int cell = blockIdx.x / a.nvar, var = blockIdx.x - a.nvar * cell;
  if (a.need && !a.need[cell]) return;         // several ranks: only the cells this rank evaluates splines of
  double *s_gd = sm;                           // [nlon][nlat]
  double *s_cs = s_gd + (size_t)nlon * nlat;   // [m0][nlon][2]
  double *s_F = s_cs + (size_t)2 * m0 * nlon;  // [2][m0][nlat]   (Fc, Fs)
  double *s_P = s_F + (size_t)2 * m0 * nlat;   // [4 arrays][2 (cos, sin coefficient)][m0][M]
  const int tid = threadIdx.x;
  const double *src = a.dens + (size_t)var * a.Np + (size_t)cell * a.npc;
  for (int e = tid; e < nlon * nlat; e += blockDim.x)
    s_gd[e] = a.raw ? src[e] : src[e] / a.w[e % nlat];  // g detJ = slist g / w
  for (int e = tid; e < 2 * m0 * nlon; e += blockDim.x) s_cs[e] = a.cs[e];
  __syncthreads();
  // phi analysis (rfft convention): F_m(i) = sum_j gd(j,i) (cos - i sin)(m phi_j)
  for (int e = tid; e < m0 * nlat; e += blockDim.x) {
    const int m = e / nlat, i = e - m * nlat;
    double fc = 0, fs = 0;
    for (int j = 0; j < nlon; j++) {
      const double g = s_gd[j * nlat + i];
      fc = fma(g, s_cs[(m * nlon + j) * 2], fc);
      fs = fma(-g, s_cs[(m * nlon + j) * 2 + 1], fs);
    }
    s_F[m * nlat + i] = fc;
    s_F[(m0 + m) * nlat + i] = fs;
  }
  __syncthreads();
  // theta operators + the i m factors of the phi derivative, irfft weights 1/nlon (m = 0) and 2/nlon
  for (int e = tid; e < m0 * M; e += blockDim.x) {
    const int m = e / M, I = e - m * M;
    const double *r0 = a.M0 + (size_t)m * nlat * M + I, *r1 = a.M1 + (size_t)m * nlat * M + I;
    double uc = 0, us = 0, vc = 0, vs = 0;
    for (int i = 0; i < nlat; i++) {
      const double fc = s_F[m * nlat + i], fs = s_F[(m0 + m) * nlat + i];
      const double c0 = __ldg(r0 + (size_t)i * M), c1 = __ldg(r1 + (size_t)i * M);
      uc = fma(c0, fc, uc);
      us = fma(c0, fs, us);
      vc = fma(c1, fc, vc);
      vs = fma(c1, fs, vs);
    }
    const double wm = (m == 0 ? 1.0 : 2.0) / (double)nlon, dm = (double)m;
    // x(j) = sum_m wm (Fc cos - Fs sin);  d/dphi: (Fc + i Fs) i m = -m Fs + i m Fc
    double *P = s_P + (size_t)m * M + I;
    const size_t st = (size_t)m0 * M;
    P[0 * st] = wm * uc;        // u   cos
    P[1 * st] = -wm * us;       // u   sin
    P[2 * st] = wm * vc;        // u1  cos
    P[3 * st] = -wm * vs;       // u1  sin
    P[4 * st] = -wm * dm * us;  // u2  cos
    P[5 * st] = -wm * dm * uc;  // u2  sin
    P[6 * st] = -wm * dm * vs;  // u12 cos
    P[7 * st] = -wm * dm * vc;  // u12 sin
  }
  __syncthreads();
  // phi synthesis of the four arrays at every node (j, I)
  const size_t plane = (size_t)M * nlon, st = (size_t)m0 * M;
  for (int e = tid; e < nlon * M; e += blockDim.x) {
    const int j = e / M, I = e - j * M;
    double r[4] = {0, 0, 0, 0};
    for (int m = 0; m < m0; m++) {
      const double cj = s_cs[(m * nlon + j) * 2], sj = s_cs[(m * nlon + j) * 2 + 1];
      const double *P = s_P + (size_t)m * M + I;
#pragma unroll
      for (int q = 0; q < 4; q++) r[q] = fma(P[(2 * q) * st], cj, fma(P[(2 * q + 1) * st], sj, r[q]));
    }
    if (a.sp_abi) {
      double *o = a.sp_abi + (size_t)cell * 4 * a.nvar * plane + (size_t)var * plane + e;  // [4][nvar][n][m]
#pragma unroll
      for (int q = 0; q < 4; q++) o[(size_t)q * a.nvar * plane] = r[q];
    }
    if (a.sp_planes) {
      // [cell][6][2 nlat][nlon], phi fastest (singular.cu: lanes = consecutive longitudes)
      double2 *o = a.sp_planes + ((size_t)cell * 6 + 2 * var) * plane + (size_t)I * nlon + j;
      o[0] = make_double2(r[0], r[1]);
      o[plane] = make_double2(r[2], r[3]);
    }
  }
}

// which: 0 = f -> spF, 1 = g -> spG (+ the plane layout of the cached singular kernel)
__global__ void k_mark_cells(int n, const int *__restrict__ cell, int *__restrict__ need) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) need[cell[i]] = 1;
}

int spline_build_density(rbc3d_ctx *c, int which) {
  Cells &C = c->cells;
  if (!C.sb_ok || C.Np == 0) return RBC3D_OK;
  const int *need = nullptr;
  if (c->prm.nranks > 1 && C.sg_cell_active.p) {
    // with several ranks the density splines are read by the singular integrals of the owned cells and by the
    // near-singular entries of this rank's targets (which may point at cells of other ranks): skip the rest
    RBC_TRY(C.sb_need.resize(C.ncell));
    CUDA_TRY(cudaMemcpyAsync(C.sb_need.p, C.sg_cell_active.p, sizeof(int) * C.ncell, cudaMemcpyDeviceToDevice, c->stream));
    for (int k = 0; k < 3; k++) {
      const TargetList &t = c->tl[k];
      if (t.valid && t.ns.n > 0)
        k_mark_cells<<<(t.ns.n + 255) / 256, 256, 0, c->stream>>>(t.ns.n, t.ns.cell.p, C.sb_need.p);
    }
    KERNEL_CHECK();
    need = C.sb_need.p;
  }
  const size_t plane = (size_t)2 * C.nlat * C.nlon;
  BuildArgs a;
  a.ncell = C.ncell;
  a.npc = C.npc;
  a.nlat = C.nlat;
  a.nlon = C.nlon;
  a.nlat0 = C.sb_nlat0;
  a.Np = C.Np;
  a.w = C.w.p;
  a.M0 = C.sb_M0.p;
  a.M1 = C.sb_M1.p;
  a.cs = C.sb_cs.p;
  a.sp_planes = nullptr;
  a.need = need;
  a.nvar = 3;
  a.raw = 0;
  if (which == 0) {
    RBC_TRY(C.spF.resize((size_t)C.ncell * 12 * plane));
    a.dens = C.f.p;
    a.sp_abi = C.spF.p;
  } else {
    RBC_TRY(C.spG.resize((size_t)C.ncell * 12 * plane));
    a.dens = C.g.p;
    a.sp_abi = C.spG.p;
    if (C.sg_cache_ok) {
      RBC_TRY(C.spGi.resize((size_t)C.ncell * 12 * plane));
      a.sp_planes = reinterpret_cast<double2 *>(C.spGi.p);
    }
  }
  const int m0 = C.sb_nlat0, M = 2 * C.nlat;
  const size_t smem = sizeof(double) * ((size_t)C.nlon * C.nlat + (size_t)2 * m0 * C.nlon + (size_t)2 * m0 * C.nlat +
                                        (size_t)8 * m0 * M);
  CUDA_TRY(cudaFuncSetAttribute(k_spline_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_spline_build<<<C.ncell * 3, 256, smem, c->stream>>>(a);
  KERNEL_CHECK();
  c->launches++;
  if (which == 1 && a.sp_planes) C.spGi_valid = true;
  if (which == 0) C.spFi_valid = false;
  return RBC3D_OK;
}

// Rbc_BuildSurfaceSource(xFlag) on the device (ModRbc.F90:736-762): splines of x, a3 and detJ from the mesh fields,
// with the same operator as the density splines (ShAnalGau + ShFilter(nlat0) + ShSynthEqu + Spline_Build_On_Sphere)
int spline_build_geometry(rbc3d_ctx *c, const double *detj_dev) {
  Cells &C = c->cells;
  if (!C.sb_ok) {
    set_error("geometry splines on the device need rbc3d_cells_enable_device_splines first");
    return RBC3D_ESTATE;
  }
  if (C.Np == 0) return RBC3D_OK;
  const size_t plane = (size_t)2 * C.nlat * C.nlon;
  RBC_TRY(C.spx.resize((size_t)C.ncell * 12 * plane));
  RBC_TRY(C.spa3.resize((size_t)C.ncell * 12 * plane));
  RBC_TRY(C.spdetj.resize((size_t)C.ncell * 4 * plane));
  BuildArgs a;
  a.ncell = C.ncell, a.npc = C.npc, a.nlat = C.nlat, a.nlon = C.nlon, a.nlat0 = C.sb_nlat0, a.Np = C.Np;
  a.w = C.w.p, a.M0 = C.sb_M0.p, a.M1 = C.sb_M1.p, a.cs = C.sb_cs.p;
  a.sp_planes = nullptr, a.need = nullptr, a.raw = 1;
  const int m0 = C.sb_nlat0, M = 2 * C.nlat;
  const size_t smem = sizeof(double) * ((size_t)C.nlon * C.nlat + (size_t)2 * m0 * C.nlon + (size_t)2 * m0 * C.nlat +
                                        (size_t)8 * m0 * M);
  CUDA_TRY(cudaFuncSetAttribute(k_spline_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const double *src[3] = {C.x.p, C.a3.p, detj_dev};
  double *dst[3] = {C.spx.p, C.spa3.p, C.spdetj.p};
  for (int k = 0; k < 3; k++) {
    a.dens = src[k];
    a.sp_abi = dst[k];
    a.nvar = k < 2 ? 3 : 1;
    k_spline_build<<<C.ncell * a.nvar, 256, smem, c->stream>>>(a);
    KERNEL_CHECK();
    c->launches++;
  }
  return RBC3D_OK;
}

}  // namespace rbc3d
