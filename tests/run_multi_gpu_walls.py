"""Multi-GPU parity check of the wall paths, one process per GPU (NOT YET RUN ON A DEVICE: written after round 1's GPU
budget was spent; scripts/gpu_r2_first.sh runs the single-GPU counterparts first):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/run_multi_gpu_walls.py

Every rank loads the same minicase-like configuration (2 cells in a tube wall), owns the wall vertices of its z-slab
(SetActiveFlag, ModTargetList.F90:205-233) and a block of cells, applies operators #3 and #4 through
rbc3d_apply_collect and runs the wall no-slip solve (rbc3d_b200/noslip.py) with the rank sum on the devices; rank 0
compares with the CPU oracle on one rank (1e-10 on velocities, same iteration count)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from rbc3d_b200 import mtube, noslip, partition
    from rbc3d_b200.capi import TL_WALLS
    from rbc3d_b200.ewald import EwaldOperator
    from tests import util
    sus, W = mtube.minicase_like(nlat0=6, ntheta=32, nz=16)
    rng = np.random.default_rng(11)
    f_rand = rng.normal(size=W.f.shape)
    op = EwaldOperator(sus.Lb, device=local)
    op.attach_comm(world, rank, dist)
    op.set_suspension(sus, active=op.ownership_mask(sus, world, rank))
    W.f = f_rand.copy()
    op.set_walls(W, active=partition.zslab_active(W.x, sus.Lb, world, rank))
    op.PrepareSingIntOnWall()
    v3 = op.apply_collect(noslip.C1_WALL, noslip.C1_WALL, TL_WALLS, cells=True, walls=True)
    v4 = op.apply_collect(noslip.C1_WALL, 0.0, TL_WALLS, cells=False, walls=True)
    W.f = np.zeros_like(f_rand)
    s = noslip.WallNoSlipSolver(W, sus.Lb, *noslip.library_backend(op, mtube.VBKG, collect=True))
    f, niter, hist, slip = s.solve()
    ok = True
    if rank == 0:
        from oracle import oracle
        sus2, W2 = mtube.minicase_like(nlat0=6, ntheta=32, nz=16)
        W2.f = f_rand.copy()
        orc = oracle.Oracle(sus2.Lb).set_cells(sus2)
        orc.set_walls(W2)
        orc.prepare_sing_int_on_walls()
        tl = orc.wall_targets()
        e3 = util.rel_l2(v3, orc.apply(noslip.C1_WALL, noslip.C1_WALL, tl, cells=True, walls=True))
        e4 = util.rel_l2(v4, orc.apply(noslip.C1_WALL, 0.0, tl, cells=False, walls=True))
        W2.f = np.zeros_like(f_rand)
        from oracle import harness
        fo, no, ho, so = noslip.WallNoSlipSolver(W2, sus2.Lb, *harness.noslip_backend(orc, mtube.VBKG)).solve()
        ef = util.rel_l2(f, fo)
        print(f"multi-gpu walls {world} ranks: operator #3 err {e3:.2e}, operator #4 err {e4:.2e}, "
              f"no-slip iterations {niter} (oracle {no}), traction err {ef:.2e}")
        ok = e3 < 1e-10 and e4 < 1e-10 and niter <= no and (niter != no or ef < 1e-6)
    op.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_WALLS_OK" if ok else "MULTI_GPU_WALLS_FAIL")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
