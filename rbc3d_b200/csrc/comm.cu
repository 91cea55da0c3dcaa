// comm.cu -- multi-GPU plumbing of the operator (one context per GPU, NCCL over NVLink / NVSwitch).
//
// The reference replicates all surface state on every MPI rank, gives every rank the targets of its z-slab
// (SetActiveFlag, ModTargetList.F90:205-233) and sums the per-rank velocity arrays with TargetList_CollectArray
// (ModTargetList.F90:172-202 -> CollectArray, ModConf.F90:531-597: AllGatherV + scatter-add).  Here:
//   * targets: the caller's `active` flags select each rank's rows (the harness owns whole cells by the z-slab of their
//     centroid, so a cell's splines, caches and singular integrals live on one GPU);
//   * PME: the mesh is decomposed into z-slabs of planes as in the reference (DomainDecomp, ModConf.F90:412-437;
//     Init_PFFTW, ModPFFTW.F90:56-89).  A rank spreads every source whose B-spline support touches its planes -- no
//     mesh reduction (ModPME.F90:428-429) --, transforms its planes in (x, y), the spectra change hands z-slabs ->
//     y-slabs with grouped ncclSend / ncclRecv (the MPI_Alltoallv of ModPFFTW.F90:188-316), the transform in z and the
//     k-space multiplier run on y-slabs, and the way back mirrors it; the velocity planes a rank's targets reach into
//     beyond its slab come from their owners (Update_Buff_Vel, ModPME.F90:354-396) -- pme.cu;
//   * results: rbc3d_collect_array / the resident path sum the velocity rows with ncclAllReduce (rows of
//     inactive targets are zero), which is CollectArray's MPI_SUM semantics without the index traffic.
#include "rbc3d_internal.h"

#ifdef RBC3D_WITH_NCCL
#include <nccl.h>
#endif

namespace rbc3d {

#ifdef RBC3D_WITH_NCCL
#define NCCL_TRY(expr)                                                                       \
  do {                                                                                       \
    ncclResult_t r__ = (expr);                                                               \
    if (r__ != ncclSuccess) {                                                                \
      rbc3d::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, ncclGetErrorString(r__)); \
      return RBC3D_ECUDA;                                                                    \
    }                                                                                        \
  } while (0)
#endif

int comm_allreduce_sum(rbc3d_ctx *c, double *buf, size_t n) {
  if (c->prm.nranks <= 1 || n == 0) return RBC3D_OK;
#ifdef RBC3D_WITH_NCCL
  NCCL_TRY(ncclAllReduce(buf, buf, n, ncclDouble, ncclSum, (ncclComm_t)c->nccl_comm, c->stream));
  return RBC3D_OK;
#else
  set_error("library built without NCCL");
  return RBC3D_EINVAL;
#endif
}

// in-place all-gather: rank r's block sits at buf + r * count
int comm_allgather_inplace(rbc3d_ctx *c, double *buf, size_t count) {
  if (c->prm.nranks <= 1 || count == 0) return RBC3D_OK;
#ifdef RBC3D_WITH_NCCL
  NCCL_TRY(ncclAllGather(buf + (size_t)c->prm.rank * count, buf, count, ncclDouble, (ncclComm_t)c->nccl_comm, c->stream));
  return RBC3D_OK;
#else
  set_error("library built without NCCL");
  return RBC3D_EINVAL;
#endif
}

int comm_allgather_ints(rbc3d_ctx *c, const int *send, int *recv, size_t count) {
  if (c->prm.nranks <= 1) {
    if (send != recv) CUDA_TRY(cudaMemcpyAsync(recv, send, sizeof(int) * count, cudaMemcpyDeviceToDevice, c->stream));
    return RBC3D_OK;
  }
#ifdef RBC3D_WITH_NCCL
  NCCL_TRY(ncclAllGather(send, recv, count, ncclInt32, (ncclComm_t)c->nccl_comm, c->stream));
  return RBC3D_OK;
#else
  set_error("library built without NCCL");
  return RBC3D_EINVAL;
#endif
}

// point-to-point pieces of the slab transposes and the velocity-mesh halo exchange: every rank issues its sends and
// receives inside one group (ncclGroupStart / ncclGroupEnd), NCCL pairs them up and runs them concurrently over NVLink
int comm_group_begin() {
#ifdef RBC3D_WITH_NCCL
  NCCL_TRY(ncclGroupStart());
#endif
  return RBC3D_OK;
}
int comm_group_end() {
#ifdef RBC3D_WITH_NCCL
  NCCL_TRY(ncclGroupEnd());
#endif
  return RBC3D_OK;
}
int comm_send(rbc3d_ctx *c, const void *buf, size_t bytes, int peer) {
#ifdef RBC3D_WITH_NCCL
  if (bytes == 0) return RBC3D_OK;
  NCCL_TRY(ncclSend(buf, bytes / 8, ncclDouble, peer, (ncclComm_t)c->nccl_comm, c->stream));
  return RBC3D_OK;
#else
  set_error("library built without NCCL");
  return RBC3D_EINVAL;
#endif
}
int comm_recv(rbc3d_ctx *c, void *buf, size_t bytes, int peer) {
#ifdef RBC3D_WITH_NCCL
  if (bytes == 0) return RBC3D_OK;
  NCCL_TRY(ncclRecv(buf, bytes / 8, ncclDouble, peer, (ncclComm_t)c->nccl_comm, c->stream));
  return RBC3D_OK;
#else
  set_error("library built without NCCL");
  return RBC3D_EINVAL;
#endif
}

void comm_destroy(rbc3d_ctx *c) {
#ifdef RBC3D_WITH_NCCL
  if (c->nccl_comm) ncclCommDestroy((ncclComm_t)c->nccl_comm);
#endif
  c->nccl_comm = nullptr;
}

}  // namespace rbc3d

using namespace rbc3d;

extern "C" {

int rbc3d_comm_unique_id(void *id128) {
  if (!id128) return RBC3D_EINVAL;
#ifdef RBC3D_WITH_NCCL
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  NCCL_TRY(ncclGetUniqueId(&id));
  memcpy(id128, &id, sizeof id);
  return RBC3D_OK;
#else
  set_error("library built without NCCL");
  return RBC3D_EINVAL;
#endif
}

int rbc3d_ctx_attach_comm(rbc3d_ctx *c, int nranks, int rank, const void *id128) {
  if (!c || nranks < 1 || rank < 0 || rank >= nranks) return RBC3D_EINVAL;
  // ownership (cell blocks, wall source ownership, per-rank target lists, PME slabs) is derived from nranks / rank when
  // a list is built: anything built before the communicator is attached would keep single-rank ownership and be counted
  // nranks times by the reductions
  if (c->cells.geom_set || c->walls.geom_set || c->walls.NE > 0 || c->tl[0].valid || c->tl[1].valid || c->tl[2].valid) {
    set_error("rbc3d_ctx_attach_comm must precede rbc3d_cells_set_geometry, rbc3d_walls_set and rbc3d_targets_set_raw");
    return RBC3D_ESTATE;
  }
  if (nranks == 1) {
    c->prm.nranks = 1;
    c->prm.rank = 0;
    return RBC3D_OK;
  }
#ifdef RBC3D_WITH_NCCL
  if (!id128) return RBC3D_EINVAL;
  CUDA_TRY(cudaSetDevice(c->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  ncclComm_t comm;
  NCCL_TRY(ncclCommInitRank(&comm, nranks, id, rank));
  c->nccl_comm = comm;
  c->prm.nranks = nranks;
  c->prm.rank = rank;
  return pme_slab_setup(c);
#else
  set_error("library built without NCCL");
  return RBC3D_EINVAL;
#endif
}

// TargetList_CollectArray (ModTargetList.F90:172-202): v <- sum over ranks, host SoA(3,n)
int rbc3d_collect_array(rbc3d_ctx *c, int tlist, double *v) {
  if (!c || tlist < 0 || tlist > 2 || !v) return RBC3D_EINVAL;
  TargetList &t = c->tl[tlist];
  if (!t.valid) return RBC3D_ESTATE;
  if (c->prm.nranks <= 1 || t.n == 0) return RBC3D_OK;
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t n3 = 3 * (size_t)t.n;
  RBC_TRY(t.host_io.resize(n3));
  t_begin(c, RBC3D_T_COMM);
  CUDA_TRY(cudaMemcpyAsync(t.host_io.p, v, sizeof(double) * n3, cudaMemcpyHostToDevice, c->stream));
  RBC_TRY(comm_allreduce_sum(c, t.host_io.p, n3));
  CUDA_TRY(cudaMemcpyAsync(v, t.host_io.p, sizeof(double) * n3, cudaMemcpyDeviceToHost, c->stream));
  t_end(c, RBC3D_T_COMM);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RBC3D_OK;
}

}  // extern "C"
