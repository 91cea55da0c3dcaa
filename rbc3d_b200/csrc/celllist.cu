// celllist.cu -- GPU cell lists: cell id per point, stable sort by cell, prefix-scan offsets.
// Replaces HashTable_Index / HashTable_Build (ModHashTable.F90:23-91): the linked list (hoc/next) becomes a
// sorted index array + offsets.  Cell ids are bit-exact with the reference's floor(x*iLbNc) arithmetic; inside
// a cell points are kept in ascending index order (stable radix sort), so every summation order is deterministic.
// The same machinery bins points by PME block (blocks of PME_BLK^3 mesh cells) for spreading / interpolation.
#include <cub/cub.cuh>

#include "device_math.cuh"
#include "rbc3d_internal.h"

namespace rbc3d {

__global__ void k_cid_realspace(int n, const double *__restrict__ x, const int *__restrict__ active, Params prm,
                                int ncells, int *__restrict__ cid, int *__restrict__ count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = ncells;  // sentinel bucket: inactive points
  if (!active || active[i]) {
    int i1 = cell_coord(x[i], prm.iLbNc[0], prm.Nc[0]);
    int i2 = cell_coord(x[(size_t)n + i], prm.iLbNc[1], prm.Nc[1]);
    int i3 = cell_coord(x[2 * (size_t)n + i], prm.iLbNc[2], prm.Nc[2]);
    c = i1 + prm.Nc[0] * (i2 + prm.Nc[1] * i3);
  }
  cid[i] = c;
  atomicAdd(&count[c], 1);
}

// PME block of a point: mesh cell floor(x*Nb/Lb) (the product rounded like ModPME.F90:420-427), wrapped into
// [0,Nb), divided by the block edge.
// zfast: key = bz + nbz * (bx + nbx * by) -- the column walks of the P = 8 kernels (pme.cu) run along z.
__global__ void k_cid_pme(int n, const double *__restrict__ x, const int *__restrict__ active, Params prm,
                          int bx, int by, int bz, int nbx, int nby, int nbz, int zfast, int *__restrict__ cid,
                          int *__restrict__ count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = nbx * nby * nbz;
  if (!active || active[i]) {
    int m0 = imodulo((int)floor(__dmul_rn(x[i], prm.ih[0])), prm.Nb[0]);
    int m1 = imodulo((int)floor(__dmul_rn(x[(size_t)n + i], prm.ih[1])), prm.Nb[1]);
    int m2 = imodulo((int)floor(__dmul_rn(x[2 * (size_t)n + i], prm.ih[2])), prm.Nb[2]);
    c = zfast ? (m2 / bz) + nbz * ((m0 / bx) + nbx * (m1 / by)) : (m0 / bx) + nbx * ((m1 / by) + nby * (m2 / bz));
  }
  cid[i] = c;
  atomicAdd(&count[c], 1);
}

__global__ void k_iota(int n, int *v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}

static int sort_by_cell(rbc3d_ctx *c, CellList &cl, int n, int ncells) {
  // start = exclusive scan of the counts currently stored in cl.start[0..ncells]
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cl.start.p, cl.start.p, ncells + 2, c->stream);
  size_t sort_bytes = 0;
  int bits = 1;
  while ((1 << bits) <= ncells) bits++;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, cl.cid.p, cl.keys_tmp.p, cl.vals_tmp.p, cl.order.p, n, 0,
                                  bits, c->stream);
  size_t need = tmp_bytes > sort_bytes ? tmp_bytes : sort_bytes;
  RBC_TRY(cl.cub_tmp.resize(need + 256));
  size_t avail = cl.cub_tmp.n;
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(cl.cub_tmp.p, avail, cl.start.p, cl.start.p, ncells + 2, c->stream));
  k_iota<<<(n + 255) / 256, 256, 0, c->stream>>>(n, cl.vals_tmp.p);
  KERNEL_CHECK();
  avail = cl.cub_tmp.n;
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(cl.cub_tmp.p, avail, cl.cid.p, cl.keys_tmp.p, cl.vals_tmp.p,
                                           cl.order.p, n, 0, bits, c->stream));
  int ns = 0;
  CUDA_TRY(cudaMemcpyAsync(&ns, cl.start.p + ncells, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  cl.n = n;
  cl.n_sorted = ns;
  c->launches += 3;
  return RBC3D_OK;
}

static int alloc_list(CellList &cl, int n, int ncells) {
  RBC_TRY(cl.cid.resize(n));
  RBC_TRY(cl.order.resize(n));
  RBC_TRY(cl.keys_tmp.resize(n));
  RBC_TRY(cl.vals_tmp.resize(n));
  RBC_TRY(cl.start.resize((size_t)ncells + 2));
  return RBC3D_OK;
}

int celllist_build_realspace(rbc3d_ctx *c, CellList &cl, int n, const double *x, const int *active) {
  const Params &p = c->prm;
  int ncells = p.Nc[0] * p.Nc[1] * p.Nc[2];
  RBC_TRY(alloc_list(cl, n > 0 ? n : 1, ncells));
  CUDA_TRY(cudaMemsetAsync(cl.start.p, 0, sizeof(int) * ((size_t)ncells + 2), c->stream));
  if (n == 0) {
    cl.n = cl.n_sorted = 0;
    return RBC3D_OK;
  }
  k_cid_realspace<<<(n + 255) / 256, 256, 0, c->stream>>>(n, x, active, p, ncells, cl.cid.p, cl.start.p);
  KERNEL_CHECK();
  c->launches++;
  return sort_by_cell(c, cl, n, ncells);
}

int celllist_build_pme(rbc3d_ctx *c, CellList &cl, int n, const double *x, const int *active, const int blk[3],
                       bool zfast) {
  const Params &p = c->prm;
  int nb[3];
  for (int d = 0; d < 3; d++) nb[d] = (p.Nb[d] + blk[d] - 1) / blk[d];
  int ncells = nb[0] * nb[1] * nb[2];
  RBC_TRY(alloc_list(cl, n > 0 ? n : 1, ncells));
  CUDA_TRY(cudaMemsetAsync(cl.start.p, 0, sizeof(int) * ((size_t)ncells + 2), c->stream));
  if (n == 0) {
    cl.n = cl.n_sorted = 0;
    return RBC3D_OK;
  }
  k_cid_pme<<<(n + 255) / 256, 256, 0, c->stream>>>(n, x, active, p, blk[0], blk[1], blk[2], nb[0], nb[1], nb[2],
                                                    zfast ? 1 : 0, cl.cid.p, cl.start.p);
  KERNEL_CHECK();
  c->launches++;
  return sort_by_cell(c, cl, n, ncells);
}

// B-spline weights of the sorted points of a PME list (BSplineFunc at x*Nb/Lb, ModPME.F90:418-424), P = 8
__global__ void k_pme_weights(int ns, int n, const int *__restrict__ order, const double *__restrict__ x, Params prm,
                              double *__restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = i / 3, ax = i - 3 * s;
  if (s >= ns) return;
  const int p = order[s];
  const double u = __dmul_rn(x[(size_t)ax * n + p], prm.ih[ax]);
  int imin;
  double ww[8];
  bspline_func<8>(u, 8, imin, ww);
  double *dst = w + (size_t)s * PME_WREC + ax * 8;
  if (ax < 2) {
#pragma unroll
    for (int q = 0; q < 8; q++) dst[q] = ww[q];
  } else {
    // the z weights are stored in ring order: the walk kernels keep plane u in ring slot u & 7, so with the point in
    // z cell cz the slot d holds plane cz - 7 + ((d - cz + 7) & 7)
    const int cz = imodulo(imin + 7, prm.Nb[2]);
#pragma unroll
    for (int q = 0; q < 8; q++) dst[(q + cz + 1) & 7] = ww[q];
  }
  if (ax == 0) {
    w[(size_t)s * PME_WREC + 24] = __hiloint2double(0, imodulo(imin + 7, prm.Nb[0]));
    w[(size_t)s * PME_WREC + 25] = 0.0;
  }
}

int celllist_pme_weights(rbc3d_ctx *c, CellList &cl, const double *x) {
  const int ns = cl.n_sorted;
  RBC_TRY(cl.w.resize((size_t)(ns > 0 ? ns : 1) * PME_WREC));
  if (ns == 0) return RBC3D_OK;
  k_pme_weights<<<(3 * ns + 255) / 256, 256, 0, c->stream>>>(ns, cl.n, cl.order.p, x, c->prm, cl.w.p);
  KERNEL_CHECK();
  c->launches++;
  return RBC3D_OK;
}

int device_exclusive_scan(rbc3d_ctx *c, int *data, int n) {
  static dbuf<char> tmp;  // grow-only scratch (one context per process and device in practice)
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, data, data, n, c->stream);
  RBC_TRY(tmp.resize(bytes + 256));
  size_t avail = tmp.n;
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, avail, data, data, n, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RBC3D_OK;
}

// ---- warp tiles of the pair kernel: (cell, first sorted position), <= 32 targets each ----
__global__ void k_tile_count(int ncells, const int *__restrict__ start, int *__restrict__ ntile) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > ncells) return;
  ntile[c] = (c < ncells) ? (start[c + 1] - start[c] + 31) / 32 : 0;
}
__global__ void k_tile_fill(int ncells, const int *__restrict__ start, const int *__restrict__ toff,
                            int2 *__restrict__ tiles) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  int b = start[c], e = start[c + 1], o = toff[c];
  for (int k = b; k < e; k += 32) tiles[o++] = make_int2(c, k);
}

int tiles_build(rbc3d_ctx *c, TargetList &t) {
  const Params &p = c->prm;
  int ncells = p.Nc[0] * p.Nc[1] * p.Nc[2];
  CellList &cl = t.cl;
  dbuf<int> &tmp = cl.keys_tmp;  // reuse: [ncells+1] tile counts / offsets
  RBC_TRY(tmp.resize((size_t)ncells + 2));
  k_tile_count<<<(ncells + 256) / 256, 256, 0, c->stream>>>(ncells, cl.start.p, tmp.p);
  KERNEL_CHECK();
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, tmp.p, tmp.p, ncells + 1, c->stream);
  RBC_TRY(cl.cub_tmp.resize(bytes + 256));
  size_t avail = cl.cub_tmp.n;
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(cl.cub_tmp.p, avail, tmp.p, tmp.p, ncells + 1, c->stream));
  int nt = 0;
  CUDA_TRY(cudaMemcpyAsync(&nt, tmp.p + ncells, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  RBC_TRY(t.tiles.resize(nt > 0 ? nt : 1));
  if (nt > 0) {
    k_tile_fill<<<(ncells + 255) / 256, 256, 0, c->stream>>>(ncells, cl.start.p, tmp.p, t.tiles.p);
    KERNEL_CHECK();
  }
  t.ntiles = nt;
  c->launches += 3;
  return RBC3D_OK;
}

}  // namespace rbc3d
