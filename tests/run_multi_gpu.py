"""Multi-GPU parity check, launched one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/run_multi_gpu.py

Every rank loads the same 8-cell suspension, owns a block of cells, applies the operator through the C ABI (host
path + TargetList_CollectArray, and the resident path) and rank 0 compares the summed velocities with the CPU oracle
and with the single-GPU result (tolerance 1e-10 relative L2, BASELINE.json north_star)."""
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
    from rbc3d_b200.ewald import EwaldOperator
    from tests import util
    sus = util.small_suspension(2)
    op = EwaldOperator(sus.Lb, device=local)
    op.attach_comm(world, rank, dist)
    op.set_suspension(sus, active=op.ownership_mask(sus, world, rank))
    ok = True
    res = {}
    for name, c1, c2 in [("matvec", 0.0, util.C2_MATVEC), ("rhs", util.C1_RHS, 0.0)]:
        if name == "rhs":   # second half with the sharded density upload (upload own block + all-gather)
            op.set_replicated_density(True)
            op.SourceList_UpdateDensity(f=sus.weighted(sus.f), g=sus.weighted(sus.g), spF=sus.spF, spG=sus.spG)
        v = op.apply(c1, c2)
        op.TargetList_CollectArray(v)
        op.apply_resident(c1, c2)
        vc = op.apply_collect(c1, c2)            # operator + CollectArray, reduced on the devices
        res[name] = (v, op.get_velocity(), vc)
    if rank == 0:
        from oracle import oracle
        orc = oracle.Oracle(sus.Lb).set_cells(sus)
        for name, c1, c2 in [("matvec", 0.0, util.C2_MATVEC), ("rhs", util.C1_RHS, 0.0)]:
            ref = orc.apply_cells(c1, c2, orc.cell_targets())
            e1, e2, e3 = (util.rel_l2(res[name][k], ref) for k in range(3))
            print(f"multi-gpu {world} ranks {name}: host path err {e1:.2e}, resident path err {e2:.2e}, "
                  f"apply_collect err {e3:.2e}")
            ok = ok and e1 < 1e-10 and e2 < 1e-10 and e3 < 1e-10
    op.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_OK" if ok else "MULTI_GPU_FAIL")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
