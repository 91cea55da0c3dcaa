// solver.cu -- SURVEY.md 8(f)-1: the cell velocity solve of ModVelSolver.F90 with everything resident on the device.
//
// Solve_RBC_Vel (ModVelSolver.F90:44-135) runs PETSc's GMRES on the spherical-harmonic coefficients of the surface
// velocity; every iteration MyMatMult (:523-601) synthesises the double-layer density from the coefficients
// (Glob_Sph_Trans FOUR_TO_PHYS, :641-719), applies the boundary-integral operator #2 and analyses v + g back
// (PHYS_TO_FOUR).  Behind the drop-in boundary that costs two PCIe crossings of 24 B per point and matvec plus the
// host transforms; here the Krylov vectors, the transforms and the operator stay on the GPU and only the Hessenberg
// column (<= 31 doubles) returns to the host per iteration.
//
//   * k_sh_synth / k_sh_anal: ShSynthGau / ShAnalGau (ModSphpk.F90:74-105, 240-270) truncated to degree < nlat0 as
//     dense per-cell operators (orthonormal associated Legendre table x DFT in phi), SPHEREPACK's shsgs convention:
//     f = sum_n [ a(0,n)/2 Pbar_n^0 + sum_{m>=1} Pbar_n^m (a(m,n) cos m phi - b(m,n) sin m phi) ];
//     packing of the unknowns as in Glob_Sph_Trans: a(m,n), n = m..nlat0-1, m = 0..; then b(m,n), m = 1..; the three
//     components interleaved.
//   * GMRES with the defaults the reference leaves untouched: restart 30, classical
//     Gram-Schmidt without refinement, no preconditioner, test on the recurrence residual against rtol ||b||,
//     non-zero initial guess.  Dot products are two-stage reductions in a fixed order (deterministic).
// The host mirror of the same algorithm is rbc3d_b200/gmres.py; tests/test_gpu_gmres.py checks one against the other.
#include <cmath>
#include <cstring>
#include <functional>
#include <vector>

#include "device_math.cuh"
#include "rbc3d_internal.h"

namespace rbc3d {

struct ShArgs {
  int ncell, npc, nlat, nlon, m0, Np, dofc;
  const int *cells;    // [slot] -> cell: the cells whose coefficients this rank holds (all of them on one rank)
  const double *pb;    // [m0][m0][nlat]  Pbar(m, n, i)            (synthesis)
  const double *pbw;   // [m0][m0][nlat]  Pbar w_i 2/nlon           (analysis)
  const double *cs;    // [m0][nlon][2]   cos, sin (m phi_j)
  const int *ka, *kb;  // [m0][m0] packed index of a(m,n) / b(m,n), -1 where there is none
  const double *dsw;   // [Np] detJ * w(ilat): raw density -> source-list density
  const double *coef;  // packed coefficients (synthesis in / analysis out)
  double *coef_out;
  double *g_raw;       // SoA(3,Np) synthesised field
  double *g_src;       // SoA(3,Np) field * detJ * w (slist%g)
  const double *v;     // SoA(3,Np) operator result (analysis adds g_raw: the diagonal term of MyMatMult)
  int add_g;           // analysis: 1 = v + g_raw (MyMatMult), 0 = v + 2 vbkg / Acoef (Compute_Rhs)
  double vbkg[3];
  const double *Acell; // [ncell]
};

// one CTA per (cell, component)
__global__ void __launch_bounds__(256) k_sh_synth(ShArgs a) {
  extern __shared__ double sm[];
  const int m0 = a.m0, nlat = a.nlat, nlon = a.nlon;
  const int slot = blockIdx.x / 3, comp = blockIdx.x - 3 * slot, cell = a.cells[slot];
  double *s_a = sm, *s_b = s_a + m0 * m0, *s_A = s_b + m0 * m0, *s_B = s_A + m0 * nlat;
  const double *c = a.coef + (size_t)slot * a.dofc + comp;
  for (int e = threadIdx.x; e < m0 * m0; e += blockDim.x) {
    const int ia = a.ka[e], ib = a.kb[e];
    s_a[e] = ia >= 0 ? c[3 * ia] : 0.0;
    s_b[e] = ib >= 0 ? c[3 * ib] : 0.0;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < m0 * nlat; e += blockDim.x) {
    const int m = e / nlat, i = e - m * nlat;
    double A = 0, B = 0;
    for (int n = m; n < m0; n++) {
      const double p = a.pb[((size_t)m * m0 + n) * nlat + i];
      A = fma(p, s_a[m * m0 + n], A);
      B = fma(p, s_b[m * m0 + n], B);
    }
    s_A[e] = A;
    s_B[e] = B;
  }
  __syncthreads();
  const size_t base = (size_t)comp * a.Np + (size_t)cell * a.npc;
  for (int e = threadIdx.x; e < nlon * nlat; e += blockDim.x) {
    const int j = e / nlat, i = e - j * nlat;
    double f = 0.5 * s_A[i];
    for (int m = 1; m < m0; m++) {
      const double cj = a.cs[(m * nlon + j) * 2], sj = a.cs[(m * nlon + j) * 2 + 1];
      f = fma(s_A[m * nlat + i], cj, f);
      f = fma(-s_B[m * nlat + i], sj, f);
    }
    a.g_raw[base + e] = f;
    a.g_src[base + e] = f * a.dsw[(size_t)cell * a.npc + e];
  }
}

__global__ void __launch_bounds__(256) k_sh_anal(ShArgs a) {
  extern __shared__ double sm[];
  const int m0 = a.m0, nlat = a.nlat, nlon = a.nlon;
  const int slot = blockIdx.x / 3, comp = blockIdx.x - 3 * slot, cell = a.cells[slot];
  double *s_v = sm, *s_Fc = s_v + nlon * nlat, *s_Fs = s_Fc + m0 * nlat;
  const size_t base = (size_t)comp * a.Np + (size_t)cell * a.npc;
  const double bk = a.add_g ? 0.0 : 2.0 * a.vbkg[comp] / a.Acell[cell];  // ModVelSolver.F90:497-500
  for (int e = threadIdx.x; e < nlon * nlat; e += blockDim.x)
    s_v[e] = a.v[base + e] + (a.add_g ? a.g_raw[base + e] : bk);
  __syncthreads();
  for (int e = threadIdx.x; e < m0 * nlat; e += blockDim.x) {
    const int m = e / nlat, i = e - m * nlat;
    double fc = 0, fs = 0;
    for (int j = 0; j < nlon; j++) {
      const double v = s_v[j * nlat + i];
      fc = fma(v, a.cs[(m * nlon + j) * 2], fc);
      fs = fma(-v, a.cs[(m * nlon + j) * 2 + 1], fs);
    }
    s_Fc[e] = fc;
    s_Fs[e] = fs;
  }
  __syncthreads();
  double *out = a.coef_out + (size_t)slot * a.dofc + comp;
  for (int e = threadIdx.x; e < m0 * m0; e += blockDim.x) {
    const int ia = a.ka[e], ib = a.kb[e];
    if (ia < 0) continue;
    const int m = e / m0;
    double ca = 0, cb = 0;
    for (int i = 0; i < nlat; i++) {
      const double p = a.pbw[(size_t)e * nlat + i];
      ca = fma(p, s_Fc[m * nlat + i], ca);
      cb = fma(p, s_Fs[m * nlat + i], cb);
    }
    out[3 * ia] = ca;
    if (ib >= 0) out[3 * ib] = cb;
  }
}

// ---- vector kernels of the Krylov solver: fixed-order two-stage reductions ----
constexpr int DOT_BLOCKS = 256, DOT_THREADS = 256;

// h[i] = V_i . w, i < gridDim.y: part[i][blk] = partial sum over the block's slice, and the block that finishes a
// vector last (ticket) adds the DOT_BLOCKS partials in a fixed order - one launch, a result that does not depend on
// the order the blocks ran in
__global__ void __launch_bounds__(DOT_THREADS) k_dots(int n, const double *__restrict__ V, size_t ldv,
                                                       const double *__restrict__ w, double *__restrict__ part,
                                                       unsigned *__restrict__ ticket, double *__restrict__ h) {
  __shared__ double s[DOT_THREADS / 32];
  __shared__ bool last;
  const int i = blockIdx.y;
  const double *v = V + (size_t)i * ldv;
  double acc = 0;
  for (int e = blockIdx.x * DOT_THREADS + threadIdx.x; e < n; e += DOT_BLOCKS * DOT_THREADS) acc = fma(v[e], w[e], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int q = 0; q < DOT_THREADS / 32; q++) t += s[q];
    part[(size_t)i * DOT_BLOCKS + blockIdx.x] = t;
    __threadfence();
    last = atomicAdd(&ticket[i], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last || threadIdx.x >= 32) return;
  __threadfence();
  double t = 0;  // (blocks past the end of a short vector are not launched: their partial sums would be zeros)
  for (int q = threadIdx.x; q < (int)gridDim.x; q += 32) t += __ldcg(part + (size_t)i * DOT_BLOCKS + q);
  t = warp_sum(t);
  if (threadIdx.x == 0) {
    h[i] = t;
    ticket[i] = 0;
  }
}
// w -= sum_i h[i] V_i
__global__ void k_gs_update(int n, int k, const double *__restrict__ V, size_t ldv, const double *__restrict__ h,
                            double *__restrict__ w) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  double acc = w[e];
  for (int i = 0; i < k; i++) acc = fma(-h[i], V[(size_t)i * ldv + e], acc);
  w[e] = acc;
}
// w = b - w
__global__ void k_residual(int n, const double *__restrict__ b, double *__restrict__ w) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) w[e] = b[e] - w[e];
}
// y = x / sqrt(*h2) with the squared norm still on the device (nothing written for a zero norm: exact breakdown)
__global__ void k_normalize(int n, const double *__restrict__ h2, const double *__restrict__ x, double *__restrict__ y) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const double hn = sqrt(*h2);
  if (hn > 0.0) y[e] = (1.0 / hn) * x[e];
}
// x += sum_i y[i] V_i, i ascending (the solution update of a restart cycle), up to 32 vectors per launch
struct Coef32 {
  double y[32];
};
__global__ void k_update_x(int n, int k, const double *__restrict__ V, size_t ldv, Coef32 y, double *__restrict__ x) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  double acc = x[e];
  for (int i = 0; i < k; i++) acc = fma(y.y[i], V[(size_t)i * ldv + e], acc);
  x[e] = acc;
}
// y = alpha x (+ y if add)
__global__ void k_axpy(int n, double alpha, const double *__restrict__ x, double *__restrict__ y, int add) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  y[e] = add ? fma(alpha, x[e], y[e]) : alpha * x[e];
}

// several ranks: the double-layer density of the cells a rank owns -> [slot][3][npc] (its block of the all-gather), and
// the other ranks' blocks back into the replicated source-list density SoA(3,Np)
__global__ void k_dens_pack(int nown, int npc, int Np, const int *__restrict__ cells, const double *__restrict__ g,
                            double *__restrict__ blk) {
  const size_t total = (size_t)nown * 3 * npc;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e % npc), d = (int)((e / npc) % 3), slot = (int)(e / ((size_t)3 * npc));
    blk[e] = g[(size_t)d * Np + (size_t)cells[slot] * npc + i];
  }
}
__global__ void k_dens_unpack(int R, int me, int maxown, int npc, int Np, const int *__restrict__ own_all,
                              const double *__restrict__ all, double *__restrict__ g) {
  const size_t per = (size_t)maxown * 3 * npc, total = (size_t)R * per;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / per);
    if (r == me) continue;
    const size_t q = e - (size_t)r * per;
    const int i = (int)(q % npc), d = (int)((q / npc) % 3), slot = (int)(q / ((size_t)3 * npc));
    const int cell = own_all[(size_t)r * maxown + slot];
    if (cell >= 0) g[(size_t)d * Np + (size_t)cell * npc + i] = all[e];
  }
}

static void pbar_table(int m0, int nlat, const double *th, std::vector<double> &pb) {
  // orthonormal associated Legendre functions with the Condon-Shortley phase, int_{-1}^{1} Pbar^2 dx = 1
  pb.assign((size_t)m0 * m0 * nlat, 0.0);
  for (int i = 0; i < nlat; i++) {
    const double x = cos(th[i]), s = sin(th[i]);
    double pmm = sqrt(0.5);
    for (int m = 0; m < m0; m++) {
      if (m > 0) pmm *= -sqrt((2.0 * m + 1.0) / (2.0 * m)) * s;
      double p2 = 0.0, p1 = pmm;
      pb[((size_t)m * m0 + m) * nlat + i] = pmm;
      for (int n = m + 1; n < m0; n++) {
        const double an = sqrt((4.0 * n * n - 1.0) / ((double)n * n - (double)m * m));
        const double bn = sqrt((((double)n - 1.0) * (n - 1.0) - (double)m * m) / (4.0 * (n - 1.0) * (n - 1.0) - 1.0));
        const double p = an * (x * p1 - bn * p2);
        pb[((size_t)m * m0 + n) * nlat + i] = p;
        p2 = p1;
        p1 = p;
      }
    }
  }
}

int solver_setup(rbc3d_ctx *c, int nlat0, const double *detj_host) {
  Cells &C = c->cells;
  Solver &S = c->solver;
  S.ok = false;
  if (!C.geom_set || nlat0 < 1 || nlat0 > C.nlat) return RBC3D_EINVAL;
  if (!C.sb_ok) {
    set_error("rbc3d_solver_setup needs rbc3d_cells_enable_device_splines (the density splines are built on the device)");
    return RBC3D_ESTATE;
  }
  const int nlat = C.nlat, nlon = C.nlon, m0 = nlat0;
  S.m0 = m0;
  S.dofc = 3 * m0 * m0;
  S.dof = (size_t)C.ncell * S.dofc;
  std::vector<double> pb, pbw((size_t)m0 * m0 * nlat), cs((size_t)m0 * nlon * 2);
  pbar_table(m0, nlat, C.h_th.data(), pb);
  const double wphi = RBC_TWO_PI / nlon;
  for (int m = 0; m < m0; m++)
    for (int n = 0; n < m0; n++)
      for (int i = 0; i < nlat; i++)  // h_w = Gauss weight * 2 pi / nlon (ModRbc.F90:93-95)
        pbw[((size_t)m * m0 + n) * nlat + i] = pb[((size_t)m * m0 + n) * nlat + i] * (C.h_w[i] / wphi) * (2.0 / nlon);
  for (int m = 0; m < m0; m++)
    for (int j = 0; j < nlon; j++) {
      cs[((size_t)m * nlon + j) * 2] = cos(m * C.h_phi[j]);
      cs[((size_t)m * nlon + j) * 2 + 1] = sin(m * C.h_phi[j]);
    }
  std::vector<int> ka(m0 * m0, -1), kb(m0 * m0, -1);
  int k = 0;
  for (int m = 0; m < m0; m++)
    for (int n = m; n < m0; n++) ka[m * m0 + n] = k++;
  for (int m = 1; m < m0; m++)
    for (int n = m; n < m0; n++) kb[m * m0 + n] = k++;
  auto up = [&](dbuf<double> &d, const std::vector<double> &h) -> int {
    RBC_TRY(d.resize(h.size()));
    CUDA_TRY(cudaMemcpyAsync(d.p, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, c->stream));
    return RBC3D_OK;
  };
  RBC_TRY(up(S.pb, pb));
  RBC_TRY(up(S.pbw, pbw));
  RBC_TRY(up(S.cs, cs));
  RBC_TRY(S.ka.resize(ka.size()));
  RBC_TRY(S.kb.resize(kb.size()));
  CUDA_TRY(cudaMemcpyAsync(S.ka.p, ka.data(), sizeof(int) * ka.size(), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(S.kb.p, kb.data(), sizeof(int) * kb.size(), cudaMemcpyHostToDevice, c->stream));
  // detJ * w per point
  std::vector<double> dsw((size_t)C.Np);
  for (size_t p = 0; p < (size_t)C.Np; p++) dsw[p] = detj_host[p] * C.h_w[p % nlat];
  RBC_TRY(up(S.dsw, dsw));
  RBC_TRY(S.g_raw.resize(3 * (size_t)C.Np));
  RBC_TRY(C.g.resize(3 * (size_t)C.Np));
  // unknowns of this rank: all cells on one rank; with several ranks the cells the rank owns targets of (whole cells:
  // the coefficients, Krylov vectors and SH transforms of a cell live on one GPU, dot products are all-reduced)
  const int R = c->prm.nranks;
  std::vector<int> own;
  if (R > 1) {
    own.resize(C.sg_nactive);
    if (C.sg_nactive)
      CUDA_TRY(cudaMemcpyAsync(own.data(), C.sg_active_list.p, sizeof(int) * C.sg_nactive, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
  } else {
    own.resize(C.ncell);
    for (int i = 0; i < C.ncell; i++) own[i] = i;
  }
  S.nown = (int)own.size();
  S.dof_loc = (size_t)S.nown * S.dofc;
  RBC_TRY(S.cells.resize(own.size() > 0 ? own.size() : 1));
  if (!own.empty()) CUDA_TRY(cudaMemcpyAsync(S.cells.p, own.data(), sizeof(int) * own.size(), cudaMemcpyHostToDevice, c->stream));
  S.maxown = S.nown;
  if (R > 1) {
    // every rank's list, padded to the largest count (-1), for the unpack side of the density all-gather
    RBC_TRY(S.own_all.resize((size_t)R * (C.ncell + 1)));
    int cnt = S.nown;
    CUDA_TRY(cudaMemcpyAsync(S.own_all.p, &cnt, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    RBC_TRY(comm_allgather_ints(c, S.own_all.p, S.own_all.p + 1, 1));
    std::vector<int> counts(R);
    CUDA_TRY(cudaMemcpyAsync(counts.data(), S.own_all.p + 1, sizeof(int) * R, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    int mx = 0, tot = 0;
    for (int r = 0; r < R; r++) mx = std::max(mx, counts[r]), tot += counts[r];
    if (tot != C.ncell) {
      set_error("rbc3d_solver_setup: the ranks own %d of %d cells -- every cell must have its targets active on exactly one rank", tot,
                C.ncell);
      return RBC3D_ESTATE;
    }
    S.maxown = mx;
    std::vector<int> padded(std::max(mx, 1), -1);
    std::copy(own.begin(), own.end(), padded.begin());
    RBC_TRY(S.own_all.resize((size_t)(R + 1) * std::max(mx, 1)));
    int *mine = S.own_all.p + (size_t)R * std::max(mx, 1);
    CUDA_TRY(cudaMemcpyAsync(mine, padded.data(), sizeof(int) * padded.size(), cudaMemcpyHostToDevice, c->stream));
    RBC_TRY(comm_allgather_ints(c, mine, S.own_all.p, (size_t)std::max(mx, 1)));
    RBC_TRY(S.gpack.resize((size_t)R * std::max(mx, 1) * 3 * C.npc));
  }
  CUDA_TRY(cudaStreamSynchronize(c->stream));  // host vectors go out of scope
  S.ok = true;
  return RBC3D_OK;
}

static void sh_args(rbc3d_ctx *c, ShArgs &a) {
  Cells &C = c->cells;
  Solver &S = c->solver;
  a.ncell = C.ncell, a.npc = C.npc, a.nlat = C.nlat, a.nlon = C.nlon, a.m0 = S.m0, a.Np = C.Np, a.dofc = S.dofc;
  a.pb = S.pb.p, a.pbw = S.pbw.p, a.cs = S.cs.p, a.ka = S.ka.p, a.kb = S.kb.p, a.dsw = S.dsw.p;
  a.cells = S.cells.p;
  a.coef = nullptr, a.coef_out = nullptr, a.g_raw = S.g_raw.p, a.g_src = C.g.p, a.v = nullptr;
  a.add_g = 1, a.vbkg[0] = a.vbkg[1] = a.vbkg[2] = 0.0, a.Acell = C.A.p;
}

// several ranks: every rank synthesised the density of its own cells; the source lists are replicated (every rank sums
// over all sources within rc and spreads the sources that touch its planes), so the blocks are all-gathered over NVLink
static int solver_share_density(rbc3d_ctx *c) {
  Cells &C = c->cells;
  Solver &S = c->solver;
  const int R = c->prm.nranks;
  if (R <= 1) return RBC3D_OK;
  const size_t per = (size_t)std::max(S.maxown, 1) * 3 * C.npc;
  if (S.nown > 0)
    k_dens_pack<<<c->sm_count * 4, 256, 0, c->stream>>>(S.nown, C.npc, C.Np, S.cells.p, C.g.p, S.gpack.p + (size_t)c->prm.rank * per);
  RBC_TRY(comm_allgather_inplace(c, S.gpack.p, per));
  k_dens_unpack<<<c->sm_count * 8, 256, 0, c->stream>>>(R, c->prm.rank, std::max(S.maxown, 1), C.npc, C.Np, S.own_all.p, S.gpack.p,
                                                        C.g.p);
  KERNEL_CHECK();
  c->launches += 2;
  return RBC3D_OK;
}

// b = MyMatMult(u), device vectors of the rank's unknowns (ModVelSolver.F90:523-601, c1 = 0, c2 = -1/(4 pi))
int solver_matmult(rbc3d_ctx *c, const double *u_dev, double *b_dev) {
  Cells &C = c->cells;
  Solver &S = c->solver;
  if (!S.ok) {
    set_error("rbc3d_solver_setup has not been called for this geometry");
    return RBC3D_ESTATE;
  }
  TargetList &t = c->tl[RBC3D_TL_CELLS];
  ShArgs a;
  sh_args(c, a);
  a.coef = u_dev;
  if (S.nown > 0) {
    const size_t sm_s = sizeof(double) * (2 * (size_t)S.m0 * S.m0 + 2 * (size_t)S.m0 * C.nlat);
    CUDA_TRY(cudaFuncSetAttribute(k_sh_synth, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_s));
    k_sh_synth<<<S.nown * 3, 256, sm_s, c->stream>>>(a);  // Glob_Sph_Trans(g, u, FOUR_TO_PHYS) + SourceList_UpdateDensity
    KERNEL_CHECK();
    c->launches++;
  }
  RBC_TRY(solver_share_density(c));
  C.g_set = true;
  C.spGi_valid = false;
  // operator #2 on this rank's targets; with several ranks the rows stay where they are (no CollectArray: the analysis
  // below only reads the cells this rank owns)
  c->resident_collect = c->prm.nranks > 1 ? 0 : 1;
  const int rc = rbc3d_apply_resident(c, 0.0, -1.0 / (4.0 * RBC_PI), 1, 0, RBC3D_TL_CELLS);
  c->resident_collect = 1;
  RBC_TRY(rc);
  a.v = t.v.p;
  a.coef_out = b_dev;
  if (S.nown > 0) {
    const size_t sm_a = sizeof(double) * ((size_t)C.nlon * C.nlat + 2 * (size_t)S.m0 * C.nlat);
    CUDA_TRY(cudaFuncSetAttribute(k_sh_anal, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_a));
    k_sh_anal<<<S.nown * 3, 256, sm_a, c->stream>>>(a);   // v = v + g; Glob_Sph_Trans(v, b, PHYS_TO_FOUR)
    KERNEL_CHECK();
    c->launches++;
  }
  S.nmatvec++;
  return RBC3D_OK;
}

// rhs = Compute_Rhs (ModVelSolver.F90:455-515) for the cells alone: operator #1 (c1 = 1/(4 pi), c2 = 0) on the resident
// single-layer density, + 2 vBkg / Acoef, Glob_Sph_Trans PHYS_TO_FOUR
int solver_rhs(rbc3d_ctx *c, const double vbkg[3], int use_walls, double *rhs_dev) {
  Cells &C = c->cells;
  Solver &S = c->solver;
  if (!S.ok) return RBC3D_ESTATE;
  if (!C.f_set) {
    set_error("rbc3d_solver_rhs: no single-layer density set");
    return RBC3D_ESTATE;
  }
  TargetList &t = c->tl[RBC3D_TL_CELLS];
  c->resident_collect = c->prm.nranks > 1 ? 0 : 1;
  const int rc = rbc3d_apply_resident(c, 1.0 / (4.0 * RBC_PI), 0.0, 1, use_walls, RBC3D_TL_CELLS);
  c->resident_collect = 1;
  RBC_TRY(rc);
  if (S.nown == 0) return RBC3D_OK;
  ShArgs a;
  sh_args(c, a);
  a.v = t.v.p;
  a.coef_out = rhs_dev;
  a.add_g = 0;
  for (int d = 0; d < 3; d++) a.vbkg[d] = vbkg[d];
  const size_t sm_a = sizeof(double) * ((size_t)C.nlon * C.nlat + 2 * (size_t)S.m0 * C.nlat);
  CUDA_TRY(cudaFuncSetAttribute(k_sh_anal, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_a));
  k_sh_anal<<<S.nown * 3, 256, sm_a, c->stream>>>(a);
  KERNEL_CHECK();
  c->launches++;
  return RBC3D_OK;
}

// Krylov work space shared by the two solves (cell velocities: Solver; wall tractions: WallSolver)
struct KrylovWork {
  dbuf<double> &V, &w, &part, &h;
};

// h_dev[0..k) = V_i . w on the device (summed over the ranks for sharded unknowns); nothing crosses to the host
static int dots_dev(rbc3d_ctx *c, KrylovWork &K, size_t n, bool reduce_ranks, int k, const double *V, size_t ldv,
                    const double *w, double *h_dev, int restart) {
  unsigned *ticket = reinterpret_cast<unsigned *>(K.part.p + (size_t)(restart + 2) * DOT_BLOCKS);
  const int nbx = (int)std::max<size_t>(1, std::min<size_t>(DOT_BLOCKS, (n + DOT_THREADS - 1) / DOT_THREADS));
  k_dots<<<dim3(nbx, k), DOT_THREADS, 0, c->stream>>>((int)n, V, ldv, w, K.part.p, ticket, h_dev);
  KERNEL_CHECK();
  c->launches++;
  if (reduce_ranks) RBC_TRY(comm_allreduce_sum(c, h_dev, (size_t)k));  // sharded unknowns: every rank holds its own cells'
  return RBC3D_OK;
}

// pinned landing buffer of the Hessenberg column (one solve at a time per host thread)
static double *krylov_host(int count) {
  static thread_local double *buf = nullptr;
  static thread_local int cap = 0;
  if (count > cap) {
    if (buf) cudaFreeHost(buf);
    buf = nullptr;
    cap = 0;
    if (cudaHostAlloc((void **)&buf, sizeof(double) * count, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    cap = count;
  }
  return buf;
}

// KSPGMRES, PCNONE, classical Gram-Schmidt; x_dev in: initial guess (ignored and zeroed when zero_guess: PETSc then takes
// r0 = b without a matvec), out: solution.  history[k] = residual norm after k iterations (history[0] = ||b - A x0||),
// at most maxit + 1 entries.  One iteration is matvec, projections, update, norm and normalisation queued back to back
// on the stream; only the Hessenberg column (k + 2 numbers) crosses to the host, once, at the end of the iteration.
template <class MatVec>
static int gmres_core(rbc3d_ctx *c, KrylovWork K, size_t n, bool reduce_ranks, bool zero_guess, MatVec &&matvec,
                      const double *b_dev, double *x_dev, double rtol, int restart, int maxit, int *niter, double *history,
                      bool run_ahead = false) {
  if (restart < 1 || restart > 200 || maxit < 0) return RBC3D_EINVAL;
  const int nb = (int)((n + 255) / 256 > 0 ? (n + 255) / 256 : 1);
  RBC_TRY(K.V.resize((size_t)(restart + 1) * (n > 0 ? n : 1)));
  RBC_TRY(K.w.resize(n > 0 ? n : 1));
  RBC_TRY(K.part.resize((size_t)(restart + 2) * DOT_BLOCKS + (restart + 2)));  // partial sums, then the tickets
  RBC_TRY(K.h.resize(2 * (size_t)(restart + 2)));  // two columns: the one being fetched and the one being computed
  CUDA_TRY(cudaMemsetAsync(K.part.p + (size_t)(restart + 2) * DOT_BLOCKS, 0, sizeof(double) * (restart + 2), c->stream));
  double *hcol = krylov_host(restart + 2);
  if (!hcol) return RBC3D_ENOMEM;
  // run_ahead: iteration k + 1 is queued before the column of iteration k is fetched (V_{k+1} is made on the device, so
  // nothing of iteration k + 1 waits for the host); the fetch goes through a second stream behind an event.  The GPU
  // then never idles across the host's per-iteration work, at the price of one discarded iteration when the solve
  // converges.  Only for operators whose application leaves no state a caller reads afterwards (the wall solve).
  struct Side {
    cudaStream_t st = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    ~Side() {
      for (cudaEvent_t e : ev)
        if (e) cudaEventDestroy(e);
      if (st) cudaStreamDestroy(st);
    }
  } side;
  if (run_ahead) {
    CUDA_TRY(cudaStreamCreateWithFlags(&side.st, cudaStreamNonBlocking));
    for (cudaEvent_t &e : side.ev) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  auto fetch = [&](int count) -> int {  // K.h[0..count) -> hcol
    CUDA_TRY(cudaMemcpyAsync(hcol, K.h.p, sizeof(double) * count, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return RBC3D_OK;
  };
  // iteration k of a cycle, all on the stream: w = A V_k, the projections, the update, the norm, V_{k+1}
  auto enqueue_iter = [&](int k) -> int {
    double *hd = K.h.p + (size_t)(k & 1) * (restart + 2);
    RBC_TRY(matvec(K.V.p + (size_t)k * n, K.w.p));
    RBC_TRY(dots_dev(c, K, n, reduce_ranks, k + 1, K.V.p, n, K.w.p, hd, restart));  // all projections from the same w
    k_gs_update<<<nb, 256, 0, c->stream>>>((int)n, k + 1, K.V.p, n, hd, K.w.p);
    RBC_TRY(dots_dev(c, K, n, reduce_ranks, 1, K.w.p, n, K.w.p, hd + k + 1, restart));
    k_normalize<<<nb, 256, 0, c->stream>>>((int)n, hd + k + 1, K.w.p, K.V.p + (size_t)(k + 1) * n);
    KERNEL_CHECK();
    c->launches += 2;
    if (run_ahead) CUDA_TRY(cudaEventRecord(side.ev[k & 1], c->stream));
    return RBC3D_OK;
  };
  auto fetch_iter = [&](int k) -> int {  // column of iteration k (k + 2 numbers) -> hcol
    const double *hd = K.h.p + (size_t)(k & 1) * (restart + 2);
    cudaStream_t st = run_ahead ? side.st : c->stream;
    if (run_ahead) CUDA_TRY(cudaStreamWaitEvent(side.st, side.ev[k & 1], 0));
    CUDA_TRY(cudaMemcpyAsync(hcol, hd, sizeof(double) * (k + 2), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return RBC3D_OK;
  };
  std::vector<double> H((size_t)(restart + 1) * restart), cs(restart), sn(restart), gv(restart + 1);
  RBC_TRY(dots_dev(c, K, n, reduce_ranks, 1, b_dev, n, b_dev, K.h.p, restart));
  RBC_TRY(fetch(1));
  const double ttol = fmax(rtol * sqrt(hcol[0]), 1e-50);
  int it = 0, nh = 0;
  if (zero_guess && n > 0) CUDA_TRY(cudaMemsetAsync(x_dev, 0, sizeof(double) * n, c->stream));
  for (;;) {
    if (zero_guess && it == 0) {
      CUDA_TRY(cudaMemcpyAsync(K.w.p, b_dev, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
    } else {
      // r = b - A x (KSPSetInitialGuessNonzero, and every restart: the true residual of the updated solution)
      RBC_TRY(matvec(x_dev, K.w.p));
      k_residual<<<nb, 256, 0, c->stream>>>((int)n, b_dev, K.w.p);
    }
    RBC_TRY(dots_dev(c, K, n, reduce_ranks, 1, K.w.p, n, K.w.p, K.h.p, restart));
    k_normalize<<<nb, 256, 0, c->stream>>>((int)n, K.h.p, K.w.p, K.V.p);
    RBC_TRY(fetch(1));
    const double beta = sqrt(hcol[0]);
    if (it == 0 && history) history[nh++] = beta;
    if (beta < ttol || it >= maxit) break;
    std::fill(H.begin(), H.end(), 0.0);
    std::fill(gv.begin(), gv.end(), 0.0);
    gv[0] = beta;
    int k = 0;
    double res = beta;
    RBC_TRY(enqueue_iter(0));
    while (k < restart && it < maxit) {
      const bool ahead = run_ahead && k + 1 < restart && it + 1 < maxit;
      if (ahead) RBC_TRY(enqueue_iter(k + 1));
      RBC_TRY(fetch_iter(k));
      const double hn = sqrt(hcol[k + 1]);
      auto Hm = [&](int i, int j) -> double & { return H[(size_t)i * restart + j]; };
      for (int i = 0; i <= k; i++) Hm(i, k) = hcol[i];
      Hm(k + 1, k) = hn;
      for (int i = 0; i < k; i++) {  // previous rotations
        const double t = cs[i] * Hm(i, k) + sn[i] * Hm(i + 1, k);
        Hm(i + 1, k) = -sn[i] * Hm(i, k) + cs[i] * Hm(i + 1, k);
        Hm(i, k) = t;
      }
      const double d = hypot(Hm(k, k), Hm(k + 1, k));
      cs[k] = d > 0.0 ? Hm(k, k) / d : 1.0;  // exact breakdown (a zero column): identity rotation, no NaN
      sn[k] = d > 0.0 ? Hm(k + 1, k) / d : 0.0;
      Hm(k, k) = d;
      Hm(k + 1, k) = 0.0;
      gv[k + 1] = -sn[k] * gv[k];
      gv[k] = cs[k] * gv[k];
      it++;
      k++;
      res = fabs(gv[k]);
      if (history) history[nh++] = res;
      if (res < ttol || hn == 0.0) break;  // (an iteration queued ahead is discarded: it only wrote w and V_{k+1})
      if (!ahead && k < restart && it < maxit) RBC_TRY(enqueue_iter(k));
    }
    // y = H^-1 g (upper triangular), x += V y
    std::vector<double> y(k);
    for (int i = k - 1; i >= 0; i--) {
      double s = gv[i];
      for (int j = i + 1; j < k; j++) s -= H[(size_t)i * restart + j] * y[j];
      y[i] = s / H[(size_t)i * restart + i];
    }
    for (int i0 = 0; i0 < k; i0 += 32) {
      Coef32 yc;
      const int m = std::min(32, k - i0);
      for (int i = 0; i < 32; i++) yc.y[i] = i < m ? y[i0 + i] : 0.0;
      k_update_x<<<nb, 256, 0, c->stream>>>((int)n, m, K.V.p + (size_t)i0 * n, n, yc, x_dev);
      c->launches++;
    }
    KERNEL_CHECK();
    c->launches += 2;
    if (res < ttol || it >= maxit) break;
    zero_guess = false;  // restart: true residual from the updated solution
  }
  if (niter) *niter = it;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RBC3D_OK;
}

int solver_gmres(rbc3d_ctx *c, const double *b_dev, double *x_dev, double rtol, int restart, int maxit, int *niter,
                 double *history) {
  Solver &S = c->solver;
  if (!S.ok) return RBC3D_ESTATE;
  KrylovWork K{S.V, S.w, S.part, S.h};
  return gmres_core(c, K, S.dof_loc, c->prm.nranks > 1, false,
                    [&](const double *in, double *out) { return solver_matmult(c, in, out); }, b_dev, x_dev, rtol, restart, maxit,
                    niter, history);
}

// ---------------------------------------------------------------------------------------------------------
// NoSlipWall on the device (ModNoSlip.F90:44-149): the first-kind equation for the wall tractions,
//   rhs  = -Compute_Wall_Residual_Vel   (operator #3: c1 = c2 = 1/4pi, cells + walls -> wall vertices, + vBkg; :153-195)
//   A df = MyMatMult(df)                (operator #4: c1 = 1/4pi, walls -> wall vertices; :255-308)
//   GMRES, no preconditioner, zero initial guess, rtol = eps_Ewd, at most 60 iterations (:70-87, 117); wall%f += df
// with the unknowns numbered by indxVertGlb (periodic duplicates share a number; AssembleArray :362-384: going to the
// 1-D vector the LAST duplicate wins).  Tractions, Krylov vectors and both operators stay on the device.
__global__ void k_wall_to_1d(int nindep, int NV, const int *__restrict__ last, const double *__restrict__ v, double scale,
                             double a0, double a1, double a2, double *__restrict__ u) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nindep) return;
  const int iv = last[p];
  u[3 * p + 0] = scale * (v[iv] + a0);
  u[3 * p + 1] = scale * (v[(size_t)NV + iv] + a1);
  u[3 * p + 2] = scale * (v[2 * (size_t)NV + iv] + a2);
}
__global__ void k_wall_from_1d(int NV, const int *__restrict__ indx, const double *__restrict__ u, const double *__restrict__ f0,
                               double *__restrict__ f) {
  const int iv = blockIdx.x * blockDim.x + threadIdx.x;
  if (iv >= NV) return;
  const int p = indx[iv];
#pragma unroll
  for (int d = 0; d < 3; d++) f[(size_t)d * NV + iv] = (f0 ? f0[(size_t)d * NV + iv] : 0.0) + u[3 * p + d];
}

int wall_noslip_solve(rbc3d_ctx *c, const int *indx_host, int nindep, const double vbkg[3], int use_cells, double rtol, int maxit,
                      double *f_host, int *niter, double *history, double *slip_host) {
  Walls &W = c->walls;
  WallSolver &S = c->wsolver;
  TargetList &t = c->tl[RBC3D_TL_WALLS];
  if (!W.geom_set || !t.valid || W.NV == 0) {
    set_error("rbc3d_noslip_solve: no walls set");
    return RBC3D_ESTATE;
  }
  const int NV = W.NV;
  const size_t n = 3 * (size_t)nindep;
  std::vector<int> idx0(NV), last(nindep, -1);
  for (int iv = 0; iv < NV; iv++) {
    const int p = indx_host[iv] - 1;
    if (p < 0 || p >= nindep) return RBC3D_EINVAL;
    idx0[iv] = p;
    last[p] = iv;  // vertex order: the last duplicate wins (AssembleArray, ModNoSlip.F90:373-377)
  }
  for (int p = 0; p < nindep; p++)
    if (last[p] < 0) return RBC3D_EINVAL;
  S.nindep = nindep;
  RBC_TRY(S.indx.resize(NV));
  RBC_TRY(S.last.resize(nindep));
  RBC_TRY(S.rhs.resize(n));
  RBC_TRY(S.x.resize(n));
  RBC_TRY(S.f0.resize(3 * (size_t)NV));
  RBC_TRY(S.fw.resize(3 * (size_t)NV));
  CUDA_TRY(cudaMemcpyAsync(S.indx.p, idx0.data(), sizeof(int) * NV, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(S.last.p, last.data(), sizeof(int) * nindep, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(S.f0.p, f_host, sizeof(double) * 3 * NV, cudaMemcpyHostToDevice, c->stream));
  RBC_TRY(walls_set_traction(c, S.f0.p, true));
  const double C1W = 1.0 / (4.0 * RBC_PI);  // ModNoSlip.F90:172-173, 281
  const int gi = (nindep + 255) / 256, gv = (NV + 255) / 256;
  // rhs = -(operator #3 + vBkg)
  RBC_TRY(rbc3d_apply_resident(c, C1W, C1W, use_cells, 1, RBC3D_TL_WALLS));
  k_wall_to_1d<<<gi, 256, 0, c->stream>>>(nindep, NV, S.last.p, t.v.p, -1.0, vbkg[0], vbkg[1], vbkg[2], S.rhs.p);
  KERNEL_CHECK();
  // The operator of one iteration is ~30 launches of microsecond kernels: it is captured into a CUDA graph (on the
  // second matvec, after an eager one has created every lazy buffer, plan and pair list) and replayed from then on,
  // also by later solves, for as long as nothing the graph refers to has changed.  One rank only - the collectives
  // of several ranks stay eager.  RBC3D_NOSLIP_GRAPH=0 keeps every matvec eager.
  bool try_graph = c->prm.nranks == 1;
  if (const char *e = getenv("RBC3D_NOSLIP_GRAPH")) try_graph = try_graph && atoi(e) != 0;
  const int gflags[4] = {NV, nindep, c->skip_flags, c->overlap};
  auto graph_drop = [&]() {
    if (S.gexec) cudaGraphExecDestroy(S.gexec);
    if (S.graph) cudaGraphDestroy(S.graph);
    S.gexec = nullptr;
    S.graph = nullptr;
  };
  auto graph_current = [&]() {
    return S.gexec && S.g_epoch == alloc_epoch() && S.g_geom == W.geom_version && S.g_mat == W.mat_version &&
           S.g_tl == t.version && memcmp(S.g_flags, gflags, sizeof gflags) == 0 && memcmp(&S.g_prm, &c->prm, sizeof(Params)) == 0;
  };
  if (!try_graph || !graph_current()) graph_drop();
  int ncall = 0;
  auto body = [&]() -> int {
    RBC_TRY(walls_set_traction(c, S.fw.p, true));
    return rbc3d_apply_resident(c, C1W, 0.0, 0, 1, RBC3D_TL_WALLS);
  };
  static const bool graph_dbg = getenv("RBC3D_DEBUG_GRAPH") != nullptr;
  if (graph_dbg)
    fprintf(stderr, "[noslip graph] solve: cached=%d epoch=%lld (graph %lld)\n", S.gexec != nullptr, alloc_epoch(), S.g_epoch);
  auto capture = [&]() {
    const long long l0 = c->launches;
    c->quiet = 1;
    int rc = RBC3D_ECUDA;
    if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
      rc = body();
      if (cudaStreamEndCapture(c->stream, &S.graph) != cudaSuccess || !S.graph) rc = RBC3D_ECUDA;
    }
    c->quiet = 0;
    if (rc == RBC3D_OK && cudaGraphInstantiate(&S.gexec, S.graph, 0) != cudaSuccess) rc = RBC3D_ECUDA;
    S.graph_launches = c->launches - l0;
    c->launches = l0;
    if (rc != RBC3D_OK) {  // not capturable here: stay eager
      cudaGetLastError();
      graph_drop();
      try_graph = false;
      return;
    }
    S.g_epoch = alloc_epoch();
    S.g_geom = W.geom_version;
    S.g_mat = W.mat_version;
    S.g_tl = t.version;
    memcpy(S.g_flags, gflags, sizeof gflags);
    S.g_prm = c->prm;
  };
  auto matvec = [&](const double *in, double *out) -> int {
    k_wall_from_1d<<<gv, 256, 0, c->stream>>>(NV, S.indx.p, in, nullptr, S.fw.p);  // wall%f = f, ModNoSlip.F90:273-278
    KERNEL_CHECK();
    if (try_graph && !S.gexec && ncall == 1) capture();
    if (S.gexec) {
      CUDA_TRY(cudaGraphLaunch(S.gexec, c->stream));
      c->launches += S.graph_launches;
    } else {
      RBC_TRY(body());
    }
    ncall++;
    k_wall_to_1d<<<gi, 256, 0, c->stream>>>(nindep, NV, S.last.p, t.v.p, 1.0, 0.0, 0.0, 0.0, out);
    KERNEL_CHECK();
    c->launches += 2;
    return RBC3D_OK;
  };
  KrylovWork K{S.V, S.w, S.part, S.h};
  bool run_ahead = c->prm.nranks == 1;  // RBC3D_NOSLIP_RUN_AHEAD=0: fetch every column before queueing the next iteration
  if (const char *e = getenv("RBC3D_NOSLIP_RUN_AHEAD")) run_ahead = run_ahead && atoi(e) != 0;
  RBC_TRY(gmres_core(c, K, n, false, true, matvec, S.rhs.p, S.x.p, rtol, 30, maxit, niter, history, run_ahead));
  // wall%f = f0 + df (:131-137), then the residual velocity with the new tractions (:140-146)
  k_wall_from_1d<<<gv, 256, 0, c->stream>>>(NV, S.indx.p, S.x.p, S.f0.p, S.fw.p);
  KERNEL_CHECK();
  RBC_TRY(walls_set_traction(c, S.fw.p, true));
  CUDA_TRY(cudaMemcpyAsync(f_host, S.fw.p, sizeof(double) * 3 * NV, cudaMemcpyDeviceToHost, c->stream));
  if (slip_host) {
    RBC_TRY(rbc3d_apply_resident(c, C1W, C1W, use_cells, 1, RBC3D_TL_WALLS));
    CUDA_TRY(cudaMemcpyAsync(slip_host, t.v.p, sizeof(double) * 3 * NV, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int d = 0; d < 3; d++)
      for (int iv = 0; iv < NV; iv++) slip_host[(size_t)d * NV + iv] += vbkg[d];
  }
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RBC3D_OK;
}

void solver_release(rbc3d_ctx *c) {
  Solver &S = c->solver;
  for (dbuf<double> *b : {&S.pb, &S.pbw, &S.cs, &S.dsw, &S.g_raw, &S.V, &S.w, &S.part, &S.h, &S.u, &S.b, &S.gpack}) b->release();
  S.cells.release();
  S.own_all.release();
  WallSolver &Ws = c->wsolver;
  for (dbuf<double> *b : {&Ws.V, &Ws.w, &Ws.part, &Ws.h, &Ws.rhs, &Ws.x, &Ws.f0, &Ws.fw}) b->release();
  Ws.indx.release();
  Ws.last.release();
  if (Ws.gexec) cudaGraphExecDestroy(Ws.gexec);
  if (Ws.graph) cudaGraphDestroy(Ws.graph);
  Ws.gexec = nullptr;
  Ws.graph = nullptr;
  S.ka.release();
  S.kb.release();
  S.ok = false;
}

}  // namespace rbc3d

using namespace rbc3d;

extern "C" {

int rbc3d_solver_setup(rbc3d_ctx *c, int nlat0, const double *detj) {
  if (!c || !detj) return RBC3D_EINVAL;
  CUDA_TRY(cudaSetDevice(c->device));
  return solver_setup(c, nlat0, detj);
}

int rbc3d_solver_dof(rbc3d_ctx *c, int64_t *dof) {
  if (!c || !dof || !c->solver.ok) return RBC3D_ESTATE;
  *dof = (int64_t)c->solver.dof_loc;
  return RBC3D_OK;
}

int rbc3d_solver_cells(rbc3d_ctx *c, int32_t *n, int32_t *cells, int cap) {
  if (!c || !n || !c->solver.ok) return RBC3D_ESTATE;
  CUDA_TRY(cudaSetDevice(c->device));
  *n = c->solver.nown;
  const int m = std::min(cap, c->solver.nown);
  if (cells && m > 0) CUDA_TRY(cudaMemcpy(cells, c->solver.cells.p, sizeof(int) * m, cudaMemcpyDeviceToHost));
  return RBC3D_OK;
}

int rbc3d_solver_matmult(rbc3d_ctx *c, const double *u, double *b) {
  if (!c || !u || !b) return RBC3D_EINVAL;
  if (!c->solver.ok) return RBC3D_ESTATE;
  CUDA_TRY(cudaSetDevice(c->device));
  Solver &S = c->solver;
  RBC_TRY(S.u.resize(S.dof_loc));
  RBC_TRY(S.b.resize(S.dof_loc));
  CUDA_TRY(cudaMemcpyAsync(S.u.p, u, sizeof(double) * S.dof_loc, cudaMemcpyHostToDevice, c->stream));
  RBC_TRY(solver_matmult(c, S.u.p, S.b.p));
  CUDA_TRY(cudaMemcpyAsync(b, S.b.p, sizeof(double) * S.dof_loc, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RBC3D_OK;
}

int rbc3d_solver_rhs(rbc3d_ctx *c, const double vbkg[3], int use_walls, double *rhs) {
  if (!c || !vbkg || !rhs) return RBC3D_EINVAL;
  if (!c->solver.ok) return RBC3D_ESTATE;
  CUDA_TRY(cudaSetDevice(c->device));
  Solver &S = c->solver;
  RBC_TRY(S.b.resize(S.dof_loc));
  RBC_TRY(solver_rhs(c, vbkg, use_walls, S.b.p));
  CUDA_TRY(cudaMemcpyAsync(rhs, S.b.p, sizeof(double) * S.dof_loc, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RBC3D_OK;
}

// surface velocity of the solution: Glob_Sph_Trans(v, sol, FOUR_TO_PHYS) (ModVelSolver.F90:124), host SoA(3,Np)
int rbc3d_solver_velocity(rbc3d_ctx *c, const double *sol, double *v) {
  if (!c || !sol || !v) return RBC3D_EINVAL;
  if (!c->solver.ok) return RBC3D_ESTATE;
  CUDA_TRY(cudaSetDevice(c->device));
  Cells &C = c->cells;
  Solver &S = c->solver;
  RBC_TRY(S.u.resize(S.dof_loc));
  RBC_TRY(S.w.resize(3 * (size_t)C.Np > S.dof ? 3 * (size_t)C.Np : S.dof_loc));
  CUDA_TRY(cudaMemcpyAsync(S.u.p, sol, sizeof(double) * S.dof_loc, cudaMemcpyHostToDevice, c->stream));
  ShArgs a;
  sh_args(c, a);
  a.coef = S.u.p;
  a.g_src = S.w.p;  // the weighted copy is not wanted here: scratch
  const size_t sm_s = sizeof(double) * (2 * (size_t)S.m0 * S.m0 + 2 * (size_t)S.m0 * C.nlat);
  if (S.nown > 0) {
    CUDA_TRY(cudaFuncSetAttribute(k_sh_synth, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_s));
    k_sh_synth<<<S.nown * 3, 256, sm_s, c->stream>>>(a);
    KERNEL_CHECK();
    c->launches++;
  }
  CUDA_TRY(cudaMemcpyAsync(v, S.g_raw.p, sizeof(double) * 3 * C.Np, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RBC3D_OK;
}

int rbc3d_noslip_solve(rbc3d_ctx *c, const int32_t *indx_vert_glb, int nindep, const double vbkg[3], int use_cells, double rtol,
                       int maxit, double *f, int *niter, double *history, double *slip) {
  if (!c || !indx_vert_glb || nindep < 1 || !vbkg || !f) return RBC3D_EINVAL;
  CUDA_TRY(cudaSetDevice(c->device));
  return wall_noslip_solve(c, indx_vert_glb, nindep, vbkg, use_cells, rtol, maxit, f, niter, history, slip);
}

int rbc3d_solver_gmres(rbc3d_ctx *c, const double *rhs, double *sol, double rtol, int restart, int maxit, int *niter,
                       double *history) {
  if (!c || !rhs || !sol) return RBC3D_EINVAL;
  if (!c->solver.ok) return RBC3D_ESTATE;
  CUDA_TRY(cudaSetDevice(c->device));
  Solver &S = c->solver;
  RBC_TRY(S.u.resize(S.dof_loc));
  RBC_TRY(S.b.resize(S.dof_loc));
  CUDA_TRY(cudaMemcpyAsync(S.b.p, rhs, sizeof(double) * S.dof_loc, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(S.u.p, sol, sizeof(double) * S.dof_loc, cudaMemcpyHostToDevice, c->stream));
  RBC_TRY(solver_gmres(c, S.b.p, S.u.p, rtol, restart, maxit, niter, history));
  CUDA_TRY(cudaMemcpyAsync(sol, S.u.p, sizeof(double) * S.dof_loc, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RBC3D_OK;
}

}  // extern "C"
