"""Multi-GPU check of the SHARDED cell velocity solve (rbc3d_solver_*, SURVEY.md 8(f)-1 + 8(e)), one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/run_multi_gpu_solver.py

Every rank owns the cells whose centroid lies in its z-slab and holds only their SH coefficients; MyMatMult, Compute_Rhs
and GMRES(30) run with Krylov vectors sharded over the ranks (densities all-gathered, dot products all-reduced).  Rank 0
compares with the same solve driven by the CPU oracle on one rank: same iteration count (north_star: same or fewer),
same residual history, same surface velocity."""
import copy
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
    from rbc3d_b200 import gmres, synth
    from rbc3d_b200.ewald import EwaldOperator
    from tests import util
    sus = util.small_suspension(2)
    npc = sus.nlat * sus.nlon
    op = EwaldOperator(sus.Lb, device=local)
    op.attach_comm(world, rank, dist)
    op.set_mesh(sus.ncell, sus.nlat, sus.nlon, sus.th, sus.phi, sus.w)
    op.enable_device_splines(sus.nlat0)
    act = op.ownership_mask(sus, world, rank)
    op.SourceList_UpdateCoord_mesh(sus.x, sus.a3, sus.detj, sus.Acoef, sus.Bcoef, sus.area, sus.meshSize, act)
    op.SourceList_UpdateDensity(f=sus.weighted(sus.f), g=sus.weighted(sus.g))
    op.solver_setup(sus.nlat0, sus.detj)
    own = op.solver_cells
    assert np.array_equal(own, np.flatnonzero(act.reshape(-1, npc)[:, 0])), own
    dofc = 3 * sus.nlat0 ** 2                                  # unknowns per cell (ModVelSolver.F90:66)
    assert op.solver_dof == len(own) * dofc
    # a global vector, the same on every rank; each rank passes the rows of its cells
    rng = np.random.default_rng(4)
    u_all = rng.uniform(-1, 1, (sus.ncell, dofc))
    b_loc = op.solver_matmult(np.ascontiguousarray(u_all[own]).reshape(-1))
    rhs_loc = op.solver_rhs((1.0, 0.0, 0.0))
    sol_loc, nit, hist = op.solver_gmres(rhs_loc, rtol=1e-11)
    v_loc = op.solver_velocity(sol_loc)                       # rows of this rank's cells, zeros elsewhere
    v = op.TargetList_CollectArray(v_loc.copy())
    gather = [None] * world
    dist.all_gather_object(gather, (own, b_loc.reshape(len(own), dofc), rhs_loc.reshape(len(own), dofc)))
    ok = True
    if rank == 0:
        from oracle import oracle
        b_all, rhs_all = np.zeros((sus.ncell, dofc)), np.zeros((sus.ncell, dofc))
        for o, b, r in gather:
            b_all[o], rhs_all[o] = b, r
        osus = copy.copy(sus)
        orc = oracle.Oracle(sus.Lb)

        def cpu_sl(fw):
            orc.set_cells(osus)
            return orc.apply_cells(util.C1_RHS, 0.0, orc.cell_targets())

        def cpu_dl(gw, g_raw):
            osus.g = np.ascontiguousarray(g_raw)
            synth.build_splines(osus, sus._builder, which=("G",))
            orc.set_cells(osus)
            return orc.apply_cells(0.0, util.C2_MATVEC, orc.cell_targets())

        sc = gmres.CellVelocitySolver(sus, cpu_sl, cpu_dl)
        e_b = util.rel_l2(b_all.reshape(-1), sc.matmult(u_all.reshape(-1)))
        rhs_c = sc.compute_rhs()
        e_r = util.rel_l2(rhs_all.reshape(-1), rhs_c)
        sol_c, v_c, it_c, h_c = sc.solve(rhs=rhs_c, rtol=1e-11)
        k = min(len(hist), len(h_c))
        e_v = util.rel_l2(v, v_c)
        same_hist = bool(np.allclose(hist[:k], h_c[:k], rtol=1e-6, atol=1e-9 * h_c[0]))
        print(f"sharded solver {world} ranks: matmult err {e_b:.2e}, rhs err {e_r:.2e}, GMRES {nit} its (oracle {it_c}), "
              f"history equal {same_hist}, velocity err {e_v:.2e}")
        ok = e_b < 1e-10 and e_r < 1e-10 and 0 < nit <= it_c and same_hist and e_v < 1e-8
    op.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_SOLVER_OK" if ok else "MULTI_GPU_SOLVER_FAIL")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
