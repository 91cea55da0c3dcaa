#!/usr/bin/env python
"""bench.py -- Ewald boundary-integral matvecs/s on the synthetic periodic suspension of BASELINE.json configs[2].

One "step" = one application of the cell double-layer operator that every GMRES iteration of ModVelSolver
evaluates (MyMatMult, ModVelSolver.F90:523-601: c1 = 0, c2 = -1/(4 pi), cell sources -> cell targets):
SourceList_UpdateDensity(g) + AddIntOnRbcs (pair sum, singular, near-singular, linear term) + PME_Distrib_Source +
PME_Transform + PME_Add_Interp_Vel [+ TargetList_CollectArray when N > 1].

    python bench.py [--gpus N] [--steps K] [--warmup W] [--cells 4096] [--impl reference]

* value  = matvecs/s with g / spline(g) already resident in HBM (rbc3d_apply_resident), device time (CUDA events
           recorded by the library on its own stream), max over ranks.
* e2e    = the same metric through the reference-facing C ABI with HOST buffers: rbc3d_cells_set_density(g, spG)
           + rbc3d_apply(v) -- host->device copies of g, spline(g*detJ), v and the device->host copy of v are
           inside the timed region (wall clock around the synchronous calls).
* roofline / kernels = per-kernel algorithmic work / CUDA-event time / measured peak.
* cpu_baseline = the CPU restatement of the reference operator (oracle/, "port": the Fortran reference cannot be
           built in this image) on the host cores of the GPU box, on a bounded sample of the same workload.
* --impl reference = that CPU operator alone (no GPU), same metric/config.

Inputs are larger than L2 (>= 8.8 GB of spline data + 1.2 GB of meshes at 4096 cells), so no explicit L2 flush.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

C2_MATVEC = -1.0 / (4.0 * np.pi)
GMRES_ITS_ASSUMED = 22     # iterations of the cell velocity solve at rtol 1e-11 (tests/test_gpu_gmres.py, 8 cells)
C1_RHS = 1.0 / (4.0 * np.pi)
MTUBE_WARM_STEPS = 5       # mtube block: untimed steps until the wall-GMRES iteration count settles (~21 per step)
FLOPS_DL_PAIR = 54.0      # flops per in-range double-layer pair (FMA = 2)
FLOPS_SPLINE3 = 155.0     # one bicubic interpolation of 3 variables
FLOPS_PATCH_EXTRA = 45.0  # kernel evaluation at a patch point
FLOPS_PATCH_CACHED = 110.0  # cached matvec path: one 3-variable bicubic (96) + 7 FMAs per patch point


PARTITION_TEXT = ("whole cells by the z-slab of their centroid (targets, singular/pair work, geometry caches); densities "
                  "uploaded 1/world per rank + ncclAllGather; PME by z-slabs of mesh planes (DomainDecomp): slab-local "
                  "spreading without a mesh reduction, 2-D FFT per plane, all-to-all transpose (grouped ncclSend/Recv), "
                  "1-D FFT in z + k-space multiplier on y-slabs, and back; velocity halo planes from their owners")


def n_side_of(cells: int) -> int:
    n = round(cells ** (1.0 / 3.0))
    if n ** 3 != cells:
        raise SystemExit(f"--cells must be a cube (64, 512, 4096); got {cells}")
    return n


def workload_name(cells, Nb):
    return f"synthetic periodic box, {cells} RBCs, SH order 12 (36x72 pts/cell), PME grid {Nb[0]}x{Nb[1]}x{Nb[2]}"


class ClockSampler(threading.Thread):
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([s.strip() for s in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d.get("hbm_gbs", 6650.0)), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


# ---------------------------------------------------------------------------------------------------------
def cpu_operator_sample(sus, sample_cells: int, spread_stride: int = 1, threads: int | None = None):
    """Time the CPU restatement on a bounded sample of the workload: ALL cells are sources of the real-space sum,
    the PME transform runs in full, the PME spread covers every ``spread_stride``-th cell (linear in sources),
    real-space sums and interpolation only the targets of the first ``sample_cells`` cells (exactly how a
    reference MPI rank works on its slab); both are extrapolated linearly.
    Returns (seconds per full matvec, detail dict)."""
    from oracle import oracle  # the one place bench.py runs oracle/: cpu_baseline and --impl reference
    oracle.lib().orc_set_num_threads(int(threads or host_threads()))   # torchrun exports OMP_NUM_THREADS=1
    cores = int(oracle.lib().orc_num_threads())
    orc = oracle.Oracle(sus.Lb).set_cells(sus)
    npc = sus.nlat * sus.nlon
    ns = min(sample_cells, sus.ncell)

    keep = {}

    def realspace(ncells_sample):
        act = np.zeros(sus.npoint, np.int32)
        # a compact block of cells (neighbouring lattice sites) so that the sample sees typical neighbourhoods
        act[:ncells_sample * npc] = 1
        tl_ = orc.cell_targets(active=act)
        t0_ = time.perf_counter()
        keep["v"] = orc.add_int_on_rbcs(0.0, C2_MATVEC, tl_)
        return time.perf_counter() - t0_, tl_

    # two sample sizes: the call has a cost that does not depend on the number of targets (cell list of all sources),
    # which must not be multiplied by the extrapolation factor
    ns_small = max(1, ns // 4)
    t_small, _ = realspace(ns_small) if ns_small < ns else (0.0, None)
    t_real, tl = realspace(ns)
    per_cell = (t_real - t_small) / (ns - ns_small) if ns_small < ns else t_real / ns
    t_fixed = max(0.0, t_real - per_cell * ns)
    t0 = time.perf_counter()
    gw, Bp = sus.weighted(sus.g), np.repeat(sus.Bcoef, npc)
    sel = slice(None)
    if spread_stride > 1:
        sel = (np.arange(sus.npoint) // npc) % spread_stride == 0
    xs, gs, a3s, Bs = (np.ascontiguousarray(a[..., sel]) for a in (sus.x, gw, sus.a3, Bp))
    t0 = time.perf_counter()
    orc.pme_distrib(0.0, C2_MATVEC, xs, None, gs, a3s, Bs)
    t_spread = (time.perf_counter() - t0) * (sus.npoint / xs.shape[1])
    t0 = time.perf_counter()
    orc.pme_transform()
    t_fft = time.perf_counter() - t0
    t0 = time.perf_counter()
    orc.pme_interp(tl, keep["v"])      # keep["v"]: the complete operator at the sample's targets (when the spread was full)
    t_interp = time.perf_counter() - t0
    scale = sus.ncell / ns
    total = t_fixed + per_cell * sus.ncell + t_interp * scale + t_spread + t_fft
    detail = {"cores": cores, "sample_cells": ns, "sample_cells_small": ns_small, "spread_stride": spread_stride,
              "t_realspace_sample_s": t_real, "t_realspace_small_sample_s": t_small,
              "t_realspace_fixed_s": t_fixed, "t_realspace_per_cell_s": per_cell, "t_interp_sample_s": t_interp,
              "t_spread_full_s": t_spread, "t_transform_full_s": t_fft, "extrapolated_matvec_s": total}
    if spread_stride == 1:
        detail["_v_sample"] = keep["v"][:, :ns * npc].copy()     # popped by the callers before the JSON line
    return total, detail


def sample_text(d, ncell):
    cpu_s = (d["t_realspace_sample_s"] + d.get("t_realspace_small_sample_s", 0.0) + d["t_interp_sample_s"] +
             d["t_spread_full_s"] / d["spread_stride"] + d["t_transform_full_s"])
    return (f"all {ncell} cells as real-space sources, PME FFT+scaling in full, PME spread of every "
            f"{d['spread_stride']}-th cell (x{d['spread_stride']}), real-space sums + interpolation for the targets of "
            f"{d.get('sample_cells_small', 0)} and {d['sample_cells']} cells (linear fit: fixed cost + per-cell cost x "
            f"{ncell}); {cpu_s:.1f} s of CPU work per sample")


REF_CACHE_DIR = os.path.join(ROOT, "oracle", "_cache")


def ref_cache_path(cells, seed):
    return os.path.join(REF_CACHE_DIR, "ref_matvec_c%d_s%d.npy" % (cells, seed))


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def bench_config(sus, Nb, alpha, eps, P, rc, Nc, seed):
    """The workload description both arms print (the driver compares them key by key)."""
    return {"workload": workload_name(sus.ncell, Nb), "cells": sus.ncell, "points": sus.npoint, "alpha": alpha,
            "eps": eps, "P": P, "rc": rc, "Nc": [int(v) for v in Nc], "visc_ratio": 5.0, "seed": seed}


def run_reference(args):
    """The CPU arm: ONE complete application of operator #2 at the full size (all targets of all cells) on the CPU
    restatement of the reference (oracle/, "port": the Fortran + MPI + PETSc + FFTW reference cannot be built in this
    image), every host thread, timed per stage with the wall clock.  Nothing is extrapolated: ``steps`` is the number of
    applications actually run (1; 0 warm-up) and ``ms_per_step`` their measured time.  The product library is not loaded
    in this process (parameters come from the oracle's own SetEwaldPrms).  The result is kept under oracle/_cache/ so
    that the GPU arm that follows on the same box can check ALL its targets against it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle
    from rbc3d_b200 import synth          # NumPy only: does not load librbc3d_b200.so
    oracle.build()
    threads = host_threads()              # torchrun exports OMP_NUM_THREADS=1: set the team size explicitly
    oracle.lib().orc_set_num_threads(int(threads))
    cores = int(oracle.lib().orc_num_threads())
    n_side = n_side_of(args.cells)
    sus = synth.make_suspension(n_side, seed=args.seed, with_f=False)
    orc = oracle.Oracle(sus.Lb).set_cells(sus)
    npc = sus.nlat * sus.nlon
    stage = {}

    def timed(name, fn):
        t0 = time.perf_counter()
        r = fn()
        stage[name] = time.perf_counter() - t0
        return r

    tl = orc.cell_targets()
    wall0 = time.perf_counter()
    v = timed("add_int_on_rbcs(cell list, pairs, singular, near-singular, linear)",
              lambda: orc.add_int_on_rbcs(0.0, C2_MATVEC, tl))
    gw, Bp = sus.weighted(sus.g), np.repeat(sus.Bcoef, npc)
    timed("pme_distrib_source", lambda: orc.pme_distrib(0.0, C2_MATVEC, sus.x, None, gw, sus.a3, Bp))
    timed("pme_transform(9 fwd + 3 inv FFT, scaling)", orc.pme_transform)
    timed("pme_add_interp_vel", lambda: orc.pme_interp(tl, v))
    sec = time.perf_counter() - wall0
    try:
        os.makedirs(REF_CACHE_DIR, exist_ok=True)
        np.save(ref_cache_path(sus.ncell, args.seed), v)
        with open(ref_cache_path(sus.ncell, args.seed) + ".json", "w") as fh:
            json.dump({"seconds": sec, "cores": cores, "stage_s": stage}, fh)
    except Exception:
        pass
    val = 1.0 / sec
    sample = ("one complete matvec: all %d targets of all %d cells, every cell a source, PME in full; nothing "
              "extrapolated" % (sus.npoint, sus.ncell))
    cb = {"value": val, "unit": "matvecs/s", "cores": cores, "kind": "port", "sample": sample,
          "stage_s": stage}
    out = {"impl": "reference", "metric": "Ewald BI matvecs/s", "value": val, "unit": "matvecs/s",
           "n_gpus": args.gpus, "steps": 1, "warmup": 0, "steps_requested": args.steps,
           "warmup_requested": args.warmup, "ms_per_step": sec * 1e3,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": bench_config(sus, orc.Nb, orc.alpha, orc.eps, orc.P, orc.rc, orc.Nc, args.seed),
           "cpu_baseline": cb,
           "e2e": {"value": val, "unit": "matvecs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "extrapolated": False}
    if args.ref_mtube:
        out["mtube"] = cpu_mtube(args)
    print(json.dumps(out))
    return 0


def minicase_config(seed):
    """examples/minicase on the reference's own Exodus wall mesh (rbc3d_b200/data/meshes/new_cyl_D6_L13_33.e, 1328 vertices /
    2404 triangles; rbc3d_b200.cases.minicase = minit.F90 restated); the generated tube only when the fixture is missing."""
    from rbc3d_b200 import cases, mtube
    try:
        sus, W, _ = cases.minicase(cases.mesh_file("new_cyl_D6_L13_33.e"), seed=seed)
        W.f[:] = 0.0
        return sus, W, "reference Exodus mesh new_cyl_D6_L13_33.e"
    except Exception:
        sus, W = mtube.minicase_like(seed=seed)
        return sus, W, "generated tube mesh (fixture missing)"


def carotid_walls_config(seed):
    """the two wall meshes of examples/carotid_web (carotid.e + web.e fixtures); generated tubes when they are missing"""
    from rbc3d_b200 import cases, synth
    try:
        W, Lb = cases.carotid_web_walls()
        return W, Lb, "reference Exodus meshes carotid.e + web.e"
    except Exception:
        Lb = np.array([10.5, 10.5, 30.0])
        W = synth.make_walls(Lb, [dict(radius=4.9, ntheta=120, nz=120), dict(radius=4.0, ntheta=48, nz=60)], wobble=0.02,
                             seed=seed)
        return W, Lb, "generated tubes (fixtures missing)"


def cpu_mtube(args):
    """configs[0] on the CPU restatement alone: the boundary-integral work of mtube time steps (rbc3d_b200/mtube.py) on the
    oracle, all host cores (the reference's own CPU-runnable case; the GPU arm reports the same block under "mtube")."""
    try:
        from oracle import oracle
        from rbc3d_b200 import mtube
        oracle.build()
        nsteps = max(2, args.mtube_steps)
        sus, W, _ = minicase_config(args.seed)
        from oracle import harness
        step = harness.OracleStep(oracle.Oracle(sus.Lb), sus, W)
        runs = [mtube.bi_timestep(step, advect=True) for _ in range(MTUBE_WARM_STEPS + nsteps)]
        cpu = runs[MTUBE_WARM_STEPS:]
        return {"bi_timesteps_per_s": nsteps / sum(r["seconds"]["total"] for r in cpu), "steps": nsteps,
                "ms_per_step": [r["seconds"]["total"] * 1e3 for r in cpu], "cores": os.cpu_count(), "kind": "port",
                "wall_gmres_iterations": [r["wall_iterations"] for r in runs],
                "workload": "examples/minicase-like: 2 RBCs in a periodic tube, %d wall vertices / %d triangles" % (W.NV, W.NE)}
    except Exception as exc:
        return {"error": str(exc)[:300]}


# ---------------------------------------------------------------------------------------------------------
def kernel_table(op, sus, ms, npairs, peaks, world=1):
    """Algorithmic work per application (DESIGN.md "Kernels") / CUDA-event time -> roofline fractions."""
    hbm_peak, fp64_peak = peaks
    N = sus.npoint / world      # every rank owns 1/world of the cells (targets, singular integrals, spreading)
    P3 = op.P ** 3
    Nx, Ny, Nz = op.Nb
    G = Nx * Ny * Nz
    M = (Nx // 2 + 1) * Ny * Nz
    npatch = op.nrad * op.nazm
    rows = []

    def add(name, t_ms, flops=None, bytes_=None, bound="fp64"):
        if t_ms <= 0:
            return
        r = {"kernel": name, "ms": t_ms, "bound": bound}
        if flops is not None:
            r["tflops"] = flops / (t_ms * 1e-3) / 1e12
            r["frac_fp64"] = r["tflops"] / fp64_peak if fp64_peak else None
        if bytes_ is not None:
            r["gbs"] = bytes_ / (t_ms * 1e-3) / 1e9
            r["frac_hbm"] = r["gbs"] / hbm_peak
        rows.append(r)

    # same-surface pairs stream (1 - mask) EA per unordered pair slot from the per-geometry cache (256-byte rows)
    pc_cells, pc_rows = op.pair_cache_info()
    # With the cache the kernel no longer executes the algorithmic flops (the table lookup, rsqrt and mask are in the
    # coefficient): what it moves is the coefficient stream, so the stage is rated against HBM (its nearer limit is the
    # shared-memory pipe); frac_fp64 stays in the row as the algorithmic figure.
    add("pair_sum(DL, %d of %d cells from the coefficient cache)" % (pc_cells, max(1, sus.ncell // world)), ms["pair"],
        flops=npairs * FLOPS_DL_PAIR, bytes_=(pc_rows * 256.0) if pc_rows else None, bound="hbm" if pc_rows else "fp64")
    # the cache holds (and the kernel streams) only patch points with a non-zero quadrature weight
    sg_on, sg_pts = op.sing_cache_info()
    add("singular(DL, cached geometry, %d of %d patch points)" % (sg_pts, npatch), ms["sing"],
        flops=N * npatch * FLOPS_PATCH_CACHED, bytes_=N * (sg_pts if sg_on else npatch) * 32.0, bound="hbm")
    add("near_singular", ms["nearsing"])
    add("spread(DL,6 sym comps)", ms["spread"], flops=N * P3 * (2 + 2 * 6), bytes_=80.0 * N + 2 * 6 * 8.0 * G)
    add("slab_transposes+halo(NCCL)", ms["comm"], bound="nvlink")
    add("fft_fwd(cuFFT D2Z x6)", ms["fft"], bytes_=6 * 2 * (8.0 * G + 16.0 * M), bound="hbm")
    add("kspace_scale", ms["kspace"], bytes_=(6 + 3) * 16.0 * M, bound="hbm")
    add("fft_inv(cuFFT Z2D x3)", ms["fft_inv"], bytes_=3 * 2 * (8.0 * G + 16.0 * M), bound="hbm")
    add("interp", ms["interp"], flops=N * P3 * 8.0, bytes_=3 * 8.0 * G + 56.0 * N)
    # SourceList_UpdateDensity gather + (device splines) spline(g detJ): 2 x 4 x 3 x 72 x 72 doubles written per cell
    add("density(gather + spline build)", ms["density"],
        bytes_=(3 * 8 * 2 + 4 + 8) * N + 2 * 12 * 8.0 * (2 * sus.nlat) * sus.nlon * (sus.ncell / world), bound="hbm")
    add("combine", ms["combine"], bytes_=(3 * 8 * 2 + 12) * N, bound="hbm")
    return rows


def run_gpu(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from rbc3d_b200 import capi, synth
    from rbc3d_b200.ewald import EwaldOperator, measure_fp64_peak

    n_side = n_side_of(args.cells)
    t0 = time.perf_counter()
    sus = synth.make_suspension(n_side, seed=args.seed, with_f=False)
    t_synth = time.perf_counter() - t0
    op = EwaldOperator(sus.Lb, device=local, nranks=1)
    if world > 1:
        op.attach_comm(world, rank, dist)
        op.set_replicated_density(True)   # g is replicated on the ranks, as in the reference: PCIe carries 1/world of it
    t0 = time.perf_counter()
    active = op.ownership_mask(sus, world, rank) if world > 1 else None
    op.set_suspension(sus, active=active, with_f=False)
    if not args.host_splines:
        # Rbc_BuildSurfaceSource(gFlag) of every matvec (ModVelSolver.F90:563) on the GPU: only g crosses PCIe
        op.enable_device_splines(sus.nlat0)
    t_setup = time.perf_counter() - t0
    N = sus.npoint
    g_host = np.ascontiguousarray(sus.weighted(sus.g))
    spG_host = np.ascontiguousarray(sus.spG) if args.host_splines else None
    v_host = np.zeros((3, N))
    lib = capi.load()
    for a in (g_host, spG_host, v_host):
        if a is None:
            continue
        capi.check(lib.rbc3d_host_register(a.ctypes.data, a.nbytes), "rbc3d_host_register")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident timing -------------------------------------------------------------------------------
    for _ in range(args.warmup):
        op.apply_resident(0.0, C2_MATVEC)
    l0 = op.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    barrier()
    stage_ms = {k: 0.0 for k in capi.STAGES}
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        op.apply_resident(0.0, C2_MATVEC)
        t = op.timings()
        for k in stage_ms:
            stage_ms[k] += t[k]
    barrier()
    wall = time.perf_counter() - wall0
    launches = op.launch_count() - l0
    dev_ms = stage_ms["total"] / args.steps
    ms_rank = torch.tensor([dev_ms, wall * 1e3 / args.steps], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(ms_rank, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(ms_rank[0]), float(ms_rank[1])
    for k in stage_ms:
        stage_ms[k] /= args.steps
    # One rank: the timed steps above ran with the PME chain on its own stream beside the real-space kernels (the
    # library's default for large lists), so their per-stage event times include waiting for SMs.  A second, short
    # pass with one stream gives every kernel's own duration for `kernels` / `roofline`; `value`, `ms_per_step` and
    # `critical_paths_ms` stay those of the default mode.  (Several ranks: the stage times are the overlapped ones and
    # are flagged as such.)
    stage_mode = "as timed (one stream)" if world == 1 else "as timed (two streams, PME stages flagged `overlapped`)"
    overlapped_ms = None
    if world == 1 and stage_ms.get("pme_chain", 0.0) > 0.0:
        overlapped_ms = {k: stage_ms[k] for k in ("total", "pme_chain", "real_chain")}
        nser = 1 if args.profile else max(1, min(3, args.steps))
        op.set_overlap(0)
        op.apply_resident(0.0, C2_MATVEC)
        ser = {k: 0.0 for k in capi.STAGES}
        for _ in range(nser):
            op.apply_resident(0.0, C2_MATVEC)
            t = op.timings()
            for k in ser:
                ser[k] += t[k] / nser
        op.set_overlap(-1)
        stage_ms = ser
        stage_mode = ("separate pass of %d steps on ONE stream (rbc3d_set_overlap(0)): each kernel's own duration; the "
                      "timed steps ran with the PME chain on a second stream (total %.2f ms, chains %.2f / %.2f ms)"
                      % (nser, overlapped_ms["total"], overlapped_ms["real_chain"], overlapped_ms["pme_chain"]))

    # ---- end to end through the C ABI with host buffers ---------------------------------------------------
    # The call a user makes per GMRES iteration is MyMatMult (ModVelSolver.F90:523-601): packed SH coefficients in,
    # packed SH coefficients out.  rbc3d_solver_matmult is that callback with HOST vectors: the timed region holds the
    # host->device copy of u, Glob_Sph_Trans both ways, the density splines, operator #2 and the device->host copy of b.
    # With several ranks every rank passes the coefficients of its own cells (sharded unknowns).  The older point-density
    # path (g in, v out: SourceList_UpdateDensity + operator + CollectArray) is timed next to it on one rank.
    e2e_steps = 1 if args.profile else max(1, min(args.steps, args.e2e_steps))
    solver_e2e = not args.host_splines
    e2e_point_s = None
    if solver_e2e:
        op.solver_setup(sus.nlat0, sus.detj)
        u_loc = np.random.default_rng(args.seed + rank).uniform(-1.0, 1.0, op.solver_dof)
        b_loc = np.zeros(op.solver_dof)
        for a in (u_loc, b_loc):
            if a.nbytes:
                capi.check(lib.rbc3d_host_register(a.ctypes.data, a.nbytes), "rbc3d_host_register")
        for _ in range(0 if args.profile else min(2, args.warmup)):
            op.solver_matmult(u_loc, out=b_loc)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            op.solver_matmult(u_loc, out=b_loc)
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        dof_t = torch.tensor([float(op.solver_dof)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(dof_t)
        h2d_e2e = d2h_e2e = int(dof_t[0]) * 8
    if world == 1 or not solver_e2e:
        for _ in range(0 if args.profile else min(2, args.warmup)):
            op.SourceList_UpdateDensity(g=g_host, spG=spG_host)
            op.apply_collect(0.0, C2_MATVEC, v=v_host)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            op.SourceList_UpdateDensity(g=g_host, spG=spG_host)
            op.apply_collect(0.0, C2_MATVEC, v=v_host)        # "v = 0" + operator + CollectArray (ModVelSolver.F90:571-584)
        barrier()
        e2e_point_s = (time.perf_counter() - t0) / e2e_steps
        if not solver_e2e:
            e2e_s = e2e_point_s
            h2d_e2e = g_host.nbytes + (spG_host.nbytes if spG_host is not None else 0)
            d2h_e2e = v_host.nbytes * world
    else:   # untimed: the operator applied to g_host, for the parity checks below
        op.SourceList_UpdateDensity(g=g_host, spG=spG_host)
        op.apply_collect(0.0, C2_MATVEC, v=v_host)
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_t[0])
    clocks = sampler.stop()
    v_e2e = v_host.copy()              # operator #2 applied to g: compared with the oracle below (full size)

    # ---- size-independent property at the full size: linearity in the density (after the timed regions) -----------
    full_size = None
    if world == 1 and not args.profile and not args.host_splines:
        try:
            npc_ = sus.nlat * sus.nlon
            g2 = np.ascontiguousarray(np.roll(g_host, 3 * npc_ + 17, axis=1)[::-1])    # another band-unlimited density
            a_, b_ = 0.5, -2.0
            op.SourceList_UpdateDensity(g=g2)
            v2 = op.apply_collect(0.0, C2_MATVEC).copy()
            op.SourceList_UpdateDensity(g=np.ascontiguousarray(a_ * g_host + b_ * g2))
            v3 = op.apply_collect(0.0, C2_MATVEC)
            lin = a_ * v_e2e + b_ * v2
            full_size = {"linearity_rel_l2": float(np.linalg.norm(v3 - lin) / np.linalg.norm(lin)),
                         "linearity": "||A(0.5 g - 2 g2) - (0.5 A g - 2 A g2)|| / ||.|| over all %d targets" % N}
            op.SourceList_UpdateDensity(g=g_host)
        except Exception as exc:
            full_size = {"error": str(exc)[:200]}

    # ---- the other half of BASELINE.json's metric: time-step rate of the boundary-integral part -----------------
    # One mtube step evaluates, on a NEW geometry: SourceList_UpdateCoord (cell lists, geometry caches), Compute_Rhs
    # (operator #1: c1 = 1/4pi, c2 = 0) and one GMRES solve (operator #2 x iterations).  Measured after the timed
    # region above with host buffers (steady state: device buffers already allocated); the iteration count is the
    # one the 8-cell parity solve takes at rtol 1e-11 (tests/test_gpu_gmres.py), stated as an assumption.
    timestep = None
    if not args.profile and not args.no_timestep and world == 1:   # single GPU only: an exception on one rank of a
        # multi-rank run would leave the others waiting in a collective
        try:
            dev_geom = not args.host_splines
            for arr in ((sus.x, sus.a3, sus.detj) if dev_geom else (sus.x, sus.a3, sus.spx, sus.spa3, sus.spdetj)):
                capi.check(lib.rbc3d_host_register(arr.ctypes.data, arr.nbytes), "rbc3d_host_register")   # pinned once
            barrier()
            t0 = time.perf_counter()
            if dev_geom:   # Rbc_BuildSurfaceSource(xFlag) on the device: x, a3, detJ cross PCIe, not their splines
                op.SourceList_UpdateCoord_mesh(sus.x, sus.a3, sus.detj, sus.Acoef, sus.Bcoef, sus.area, sus.meshSize,
                                               active)
            else:
                op.SourceList_UpdateCoord(sus.x, sus.a3, sus.Acoef, sus.Bcoef, sus.area, sus.meshSize, sus.spx,
                                          sus.spa3, sus.spdetj, active)
            barrier()
            t_geom = time.perf_counter() - t0
            op.SourceList_UpdateDensity(f=g_host)          # any band-limited density times the RHS operator
            t_rhs = []
            for _ in range(2):
                barrier()
                t0 = time.perf_counter()
                op.SourceList_UpdateDensity(f=g_host)
                op.apply_collect(C1_RHS, 0.0, v=v_host)
                barrier()
                t_rhs.append(time.perf_counter() - t0)
            # MyMatMult with the SH transforms and the Krylov vector on the device (rbc3d_solver_*):
            # only 2 x dof doubles cross PCIe here, none inside rbc3d_solver_gmres
            t_mm = float("nan")
            its, t_solve, resid = None, float("nan"), None
            if dev_geom:
                op.solver_setup(sus.nlat0, sus.detj)      # again: the geometry update above invalidated it
                u = np.random.default_rng(args.seed).uniform(-1.0, 1.0, op.solver_dof)
                op.solver_matmult(u)
                tm = []
                for _ in range(3):
                    barrier()
                    t0 = time.perf_counter()
                    op.solver_matmult(u)
                    barrier()
                    tm.append(time.perf_counter() - t0)
                t_mm = min(tm)
                # a REAL cell velocity solve at this size (Solve_RBC_Vel, ModVelSolver.F90:74-116: GMRES(30), rtol 1e-11,
                # zero initial guess = the first step of a run): its iteration count replaces any assumption
                op.SourceList_UpdateDensity(f=g_host)
                rhs = op.solver_rhs((0.0, 0.0, 8.0))
                barrier()
                t0 = time.perf_counter()
                _, its, hist = op.solver_gmres(rhs, rtol=1e-11)
                barrier()
                t_solve = time.perf_counter() - t0
                resid = float(hist[-1] / hist[0]) if len(hist) and hist[0] > 0 else None
            tt = torch.tensor([t_geom, min(t_rhs), t_mm, t_solve], dtype=torch.float64, device="cuda")
            if dist is not None:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_geom, t_rhs1, t_mm, t_solve = float(tt[0]), float(tt[1]), float(tt[2]), float(tt[3])
            its_used = its if its is not None else GMRES_ITS_ASSUMED
            timestep = {"geometry_splines": "device" if dev_geom else "host (uploaded)",
                        "geometry_update_ms": t_geom * 1e3, "rhs_operator_ms": t_rhs1 * 1e3,
                        "matvec_e2e_ms": (e2e_point_s or e2e_s) * 1e3, "gmres_iterations": its_used,
                        "gmres_iterations_measured": its is not None, "gmres_relative_residual": resid,
                        "gmres_solve_ms": t_solve * 1e3 if t_solve == t_solve else None,
                        "bi_timesteps_per_s": 1.0 / (t_geom + t_rhs1 + its_used * (e2e_point_s or e2e_s)),
                        "matmult_device_solver_ms": t_mm * 1e3,
                        "bi_timesteps_per_s_device_solver": (1.0 / (t_geom + t_rhs1 + t_solve)) if t_solve == t_solve else None,
                        "note": "boundary-integral part of one mtube step (membrane forces stay in the Fortran caller); "
                                "device_solver: rbc3d_solver_gmres, SH transforms and Krylov vectors on the GPU"}
        except Exception as exc:  # e.g. no room left for spline(f detJ) next to the caches
            timestep = {"error": str(exc)[:200]}
    h2d, d2h = h2d_e2e, d2h_e2e    # whole job, counted from the arrays copied inside the timed region

    if rank != 0:
        op.close()
        return 0

    hbm_peak, peak_src = measured_peaks()
    fp64_peak = measure_fp64_peak(local)
    cnt, _ = op.neighbor_signature()
    npairs = float(cnt.astype(np.float64).sum())
    rows = kernel_table(op, sus, stage_ms, npairs, (hbm_peak, fp64_peak), world)
    if world > 1:
        # the PME chain runs on its own stream beside the real-space kernels: its per-kernel event times include waiting
        # for SMs and are marked; the dominant kernel is taken from the real-space chain (the critical path)
        for r in rows:
            r["overlapped"] = r["kernel"].split("(")[0] in ("spread", "slab_transposes+halo", "fft_fwd", "fft_inv",
                                                            "kspace_scale", "interp")
    dom = max([r for r in rows if not r.get("overlapped")], key=lambda r: r["ms"])
    traffic = None
    try:   # dram__bytes_read + dram__bytes_write per launch of the dominant kernel from the committed ncu --set full capture
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_traffic.json")) as fh:
            tr = json.load(fh)
        traffic = tr.get("%s@%d" % (dom["kernel"].split("(")[0], sus.ncell // world))
    except Exception:
        traffic = None
    if dom.get("tflops") is not None and dom["bound"] == "fp64":
        roof = {"kernel": dom["kernel"], "bound": "fp64", "achieved": dom["tflops"], "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": dom["tflops"] / fp64_peak, "traffic": traffic,
                "peak_source": "FP64 FMA micro-benchmark run in this process (MEASURED_PEAKS.json has no FP64 entry)"}
    else:
        roof = {"kernel": dom["kernel"], "bound": "hbm", "achieved": dom.get("gbs"), "peak": hbm_peak, "unit": "GB/s",
                "frac": (dom.get("gbs") or 0.0) / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": (dom.get("gbs") or 0.0) * 1e9 * dom["ms"] * 1e-3,
                "frac_fp64": dom.get("frac_fp64")}

    # ---- parity at the benchmark's own size (the oracle as the checker, after every timed region) -----------------
    # (1) when the CPU arm (--impl reference) ran before on this box, its complete result is under oracle/_cache/:
    #     ALL targets are compared; (2) otherwise the complete operator at the targets of a sample of cells.
    cb = None
    parity = dict(full_size or {})
    ref_full = None
    try:
        pth = ref_cache_path(sus.ncell, args.seed)
        if os.path.exists(pth) and not args.profile:
            ref_full = np.load(pth, mmap_mode="r")
            if ref_full.shape != v_e2e.shape:
                ref_full = None
    except Exception:
        ref_full = None
    if ref_full is not None:
        num = den = 0.0
        worst = 0.0
        step_ = 64 * sus.nlat * sus.nlon
        for lo in range(0, N, step_):
            r_ = np.asarray(ref_full[:, lo:lo + step_])
            d_ = v_e2e[:, lo:lo + step_] - r_
            n2, d2 = float((d_ * d_).sum()), float((r_ * r_).sum())
            num, den = num + n2, den + d2
            worst = max(worst, (n2 / d2) ** 0.5 if d2 > 0 else 0.0)
        parity["oracle_rel_l2"] = (num / den) ** 0.5
        parity["oracle_rel_l2_worst_64_cell_block"] = worst
        parity["oracle_targets"] = int(N)
        parity["oracle_source"] = "complete CPU matvec of the --impl reference run on this box (oracle/_cache)"
        try:
            with open(ref_cache_path(sus.ncell, args.seed) + ".json") as fh:
                parity["oracle_full_matvec_s"] = json.load(fh).get("seconds")
        except Exception:
            pass
    need_sample = (world == 1 and not args.no_cpu_baseline) or (ref_full is None and not args.no_cpu_baseline
                                                                and not args.profile)
    if need_sample:
        ns_ = args.cpu_sample_cells if world == 1 else min(8, args.cpu_sample_cells)
        tcpu, detail = cpu_operator_sample(sus, ns_)
        vs = detail.pop("_v_sample", None)
        if vs is not None:   # the oracle as the checker at the FULL size: complete operator at the sample's targets
            ref_n = float(np.linalg.norm(vs))
            parity["oracle_sample_rel_l2"] = float(np.linalg.norm(v_e2e[:, :vs.shape[1]] - vs) / ref_n)
            parity["oracle_sample_targets"] = int(vs.shape[1])
        if world == 1:
            cb = {"value": 1.0 / tcpu, "unit": "matvecs/s", "cores": detail["cores"], "kind": "port",
                  "sample": sample_text(detail, sus.ncell), "extrapolated": True, "detail": detail}
            if "oracle_full_matvec_s" in parity and parity["oracle_full_matvec_s"]:
                cb["measured_full_matvec_s_of_reference_arm"] = parity["oracle_full_matvec_s"]
    parity["n_gpus"] = world
    parity["tolerance"] = 1e-10
    vals_ = [parity[k] for k in ("oracle_rel_l2", "oracle_sample_rel_l2", "linearity_rel_l2") if parity.get(k) is not None]
    parity["ok"] = all(v_ <= 1e-10 for v_ in vals_) if vals_ else None   # None: no oracle value available in this run

    out = {"metric": "Ewald BI matvecs/s", "value": world_value(1e3 / dev_ms), "unit": "matvecs/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms,
           "wall_ms_per_step": wall_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": dict(bench_config(sus, op.Nb, op.alpha, op.eps, op.P, op.rc, op.cell_list_dims(), args.seed),
                          in_range_pairs=npairs,
                          l2="inputs (>= 10 GB at 4096 cells) exceed the 126 MB L2; no explicit flush",
                          density_splines="host (uploaded)" if args.host_splines else "device (built from g every step)",
                          partition=PARTITION_TEXT if world > 1 else "single GPU"),
           "clocks": clocks,
           "e2e": {"value": 1.0 / e2e_s, "unit": "matvecs/s", "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                   "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "api": ("rbc3d_solver_matmult (MyMatMult: packed SH coefficients of the rank's cells in and out, host "
                           "vectors)") if solver_e2e else "rbc3d_cells_set_density + rbc3d_apply_collect (g in, v out)",
                   "point_density_path_ms": (e2e_point_s * 1e3) if e2e_point_s else None},
           "gpu_launches": int(launches),
           "stage_ms": stage_ms, "fft_ms": stage_ms["fft"] + stage_ms["fft_inv"],
           "stage_ms_mode": stage_mode,
           "critical_paths_ms": {"real_space_chain": (overlapped_ms or stage_ms).get("real_chain"),
                                 "pme_chain": (overlapped_ms or stage_ms).get("pme_chain"),
                                 "note": "the two chains run on two streams and join before the combine (several ranks; one rank: lists of 100 000 targets or more)"},
           "roofline": roof, "kernels": rows, "timestep": timestep,
           "peaks": {"hbm_gbs": hbm_peak, "hbm_source": peak_src, "fp64_tflops": fp64_peak,
                     "fp64_source": "in-process DFMA micro-benchmark"},
           "setup_s": {"synth": t_synth, "upload+geometry": t_setup}}
    if cb is not None:
        out["cpu_baseline"] = cb
    try:
        op.close()
        if world == 1 and not args.profile and not args.no_mtube:
            # the other half of the metric on BASELINE.json configs[0]; a child process, after this one released the GPU
            out["mtube"] = mtube_child(args)
            out["walls"] = child_block(args, "--walls-only", "walls")       # configs[4], wall-dominated operator
    except Exception as exc:   # nothing after the timed regions may cost the line
        out.setdefault("mtube", {"error": str(exc)[:300]})
    # parity evidence goes LAST so that it survives a truncated tail of the line
    for key in ("mtube", "walls"):
        blk = out.get(key) or {}
        pv = blk.get("parity_vs_oracle") if isinstance(blk, dict) else None
        if pv:
            parity[key] = {k: (max(v) if isinstance(v, list) and v and not isinstance(v[0], bool) else
                               (all(v) if isinstance(v, list) else v)) for k, v in pv.items()}
    out["parity"] = parity
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_mtube(args):
    """BASELINE.json configs[0] ("examples/minicase: a few RBCs in a periodic tube"): the boundary-integral work of one
    TimeInt_Euler step (rbc3d_b200/mtube.py: geometry update, Compute_Rhs with cells + wall, NoSlipWall = operator #3
    twice + operator #4 per wall-GMRES iteration) through the C ABI with host buffers, and the same steps on the CPU
    oracle on the box's host cores.  Tractions carry over from step to step as in a run.  Prints one JSON object; run
    as a child process of the main bench so that a failure here cannot cost the headline line."""
    from rbc3d_b200 import mtube
    from rbc3d_b200.capi import Rbc3dError
    from rbc3d_b200.ewald import EwaldOperator
    nsteps = max(2, args.mtube_steps)
    sus, W, mesh_src = minicase_config(args.seed)
    try:          # no torch in this child (its import costs more than the block): the library itself refuses without a GPU
        op = EwaldOperator(sus.Lb, device=int(os.environ.get("LOCAL_RANK", "0")))
    except Rbc3dError as exc:
        raise SystemExit("bench.py --mtube-only: no CUDA device; the product has no CPU path (%s)" % str(exc)[:160])
    step = mtube.LibraryStep(op, sus, W, device_noslip=not args.host_noslip)
    l0 = op.launch_count()
    warm = MTUBE_WARM_STEPS      # start-up transient of the wall tractions (3, 19, 60, 42 iterations), allocations, plans
    runs = [mtube.bi_timestep(step, advect=True) for _ in range(warm + nsteps)]
    launches = op.launch_count() - l0
    Nb = list(op.Nb)
    op.close()
    gpu = runs[warm:]
    out = {"workload": "examples/minicase: 2 RBCs (36x72 pts/cell, lambda = 1) in a periodic tube of radius 5, box "
                       "10.5x10.5x8, %s, %d vertices / %d triangles, vBkg = (0,0,8), PME grid %s"
                       % (mesh_src, W.NV, W.NE, "x".join(str(n) for n in Nb)),
           "between_steps": "cells translated by Ts = 0.0008 times their mean surface velocity (stand-in for the membrane "
                            "update, untimed); %d untimed start-up steps" % warm,
           "step": "geometry update + operator #1 (RHS, cells+wall -> cells) + NoSlipWall (operator #3 x 2 + operator #4 "
                   "per wall-GMRES iteration, rtol = eps_Ewd = 1e-3, <= 60); host buffers through the C ABI; NoSlipWall "
                   + ("with Krylov vectors on the host around per-matvec calls" if args.host_noslip else
                      "resident on the device (rbc3d_noslip_solve)"),
           "steps": nsteps,
           "bi_timesteps_per_s": nsteps / sum(r["seconds"]["total"] for r in gpu),
           "ms_per_step": [r["seconds"]["total"] * 1e3 for r in gpu],
           "ms_geometry_rhs_noslip": [[r["seconds"][k] * 1e3 for k in ("geometry", "rhs", "noslip")] for r in gpu],
           "wall_gmres_iterations": [r["wall_iterations"] for r in runs],
           "operator_applications": [r["operator_applications"] for r in gpu],
           "gpu_launches": int(launches)}
    if not args.no_cpu_baseline:
        from oracle import oracle
        oracle.build()
        sus2, W2, _ = minicase_config(args.seed)
        from oracle import harness
        ostep = harness.OracleStep(oracle.Oracle(sus2.Lb), sus2, W2)
        cruns = [mtube.bi_timestep(ostep, advect=True) for _ in range(warm + nsteps)]
        cpu = cruns[warm:]
        rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))  # noqa: E731
        out["cpu_baseline"] = {"value": nsteps / sum(r["seconds"]["total"] for r in cpu), "unit": "timesteps/s",
                               "cores": os.cpu_count(), "kind": "port",
                               "sample": "the same %d steps in full on the CPU oracle (OpenMP, all host cores)" % nsteps,
                               "ms_per_step": [r["seconds"]["total"] * 1e3 for r in cpu],
                               "wall_gmres_iterations": [r["wall_iterations"] for r in cruns]}
        out["parity_vs_oracle"] = {"rel_l2_cell_velocity": [rel(a["v_cells"], b["v_cells"]) for a, b in zip(runs, cruns)],
                                   "rel_l2_wall_traction": [rel(a["f_wall"], b["f_wall"]) for a, b in zip(runs, cruns)],
                                   "same_iterations": [a["wall_iterations"] == b["wall_iterations"]
                                                       for a, b in zip(runs, cruns)]}
    print(json.dumps({"mtube": out}))
    return 0


def child_block(args, flag, key):
    """Run a side block (--mtube-only / --walls-only) in a child process, after the main operator has released the GPU,
    and return its object; a failure there becomes {"error": ...} instead of costing the headline line."""
    cmd = [sys.executable, os.path.abspath(__file__), flag, "--seed", str(args.seed), "--mtube-steps",
           str(args.mtube_steps)] + (["--no-cpu-baseline"] if args.no_cpu_baseline else []) + (
               ["--host-noslip"] if getattr(args, "host_noslip", False) else [])
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "MASTER_ADDR",
                                                            "MASTER_PORT", "TORCHELASTIC_RUN_ID")}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
        for line in reversed(r.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)[key]
        return {"error": ("rc=%d " % r.returncode) + r.stderr.strip()[-300:]}
    except Exception as exc:
        return {"error": str(exc)[:300]}


def mtube_child(args):
    return child_block(args, "--mtube-only", "mtube")


def run_walls(args):
    """BASELINE.json configs[4] ("large complex wall mesh ... wall-dominated operator") at the size of
    examples/carotid_web with generated walls (17 448 vertices / 34 560 triangles in two interacting tubes, box
    10.5 x 10.5 x 30, PME grid 48 x 48 x 136; the Exodus meshes are not on the GPU box): the wall matvec of the no-slip
    solve (operator #4: c1 = 1/4pi, wall tractions -> wall vertices; self-interaction matrices + wall-wall direct loop +
    PME) through the C ABI with host buffers, the CPU oracle beside it, and their agreement.  Child process of the bench."""
    from rbc3d_b200 import synth
    from rbc3d_b200.capi import TL_WALLS, Rbc3dError
    from rbc3d_b200.ewald import EwaldOperator
    W, Lb, mesh_src = carotid_walls_config(args.seed)
    W.f = np.zeros_like(W.x)
    try:
        op = EwaldOperator(Lb, device=int(os.environ.get("LOCAL_RANK", "0")))
    except Rbc3dError as exc:
        raise SystemExit("bench.py --walls-only: no CUDA device; the product has no CPU path (%s)" % str(exc)[:160])
    t0 = time.perf_counter()
    op.set_walls(W)
    op.PrepareSingIntOnWall()
    t_prep_first = time.perf_counter() - t0                     # with the one-time costs of a new process: lazy loading
    t0 = time.perf_counter()                                    # of the kernels' code, first allocations
    op.PrepareSingIntOnWall()                                   # the same matrices again, from scratch
    t_prep = time.perf_counter() - t0
    rng = np.random.default_rng(args.seed)
    fs = [rng.normal(size=W.f.shape) for _ in range(4)]
    nrep = 5
    for f in fs[:2]:                                            # warm-up: pair lists, cuFFT plans
        op.set_wall_traction(f)
        v = op.apply(C1_RHS, 0.0, TL_WALLS, cells=False, walls=True)
    l0 = op.launch_count()
    t0 = time.perf_counter()
    for k in range(nrep * len(fs)):
        op.set_wall_traction(fs[k % len(fs)])                   # wall%f = f of every GMRES iteration (ModNoSlip.F90:273-278)
        v = op.apply(C1_RHS, 0.0, TL_WALLS, cells=False, walls=True)
    t_gpu = (time.perf_counter() - t0) / (nrep * len(fs))
    launches = op.launch_count() - l0
    rowptr, _, _ = op.wall_matrix()
    Nb = list(op.Nb)
    op.close()
    out = {"workload": "wall-dominated operator of examples/carotid_web: %s, %d + %d vertices, %d + %d triangles, box "
                       "%.2fx%.2fx%.0f, PME grid %s" % (mesh_src, W.nvert[0], W.nvert[1], W.nele[0], W.nele[1], Lb[0], Lb[1],
                                                        Lb[2], "x".join(str(n) for n in Nb)),
           "step": "operator #4 of the wall no-slip solve (set traction + AddIntOnWalls + PME, host buffers through the C ABI)",
           "wall_matvecs_per_s": 1.0 / t_gpu, "ms_per_matvec": t_gpu * 1e3, "matrix_blocks_3x3": int(rowptr[-1]),
           "prepare_sing_int_on_wall_ms": t_prep * 1e3, "set_walls_and_first_prepare_ms": t_prep_first * 1e3,
           "gpu_launches": int(launches), "steps": nrep * len(fs)}
    if not args.no_cpu_baseline:
        from oracle import oracle
        oracle.build()
        orc = oracle.Oracle(Lb)
        orc.set_walls(W, ncell=0)
        t0 = time.perf_counter()
        orc.prepare_sing_int_on_walls()
        t_prep_cpu = time.perf_counter() - t0
        tl = orc.wall_targets()
        orc.set_wall_traction(fs[0])
        orc.apply(C1_RHS, 0.0, tl, cells=False, walls=True)
        t0 = time.perf_counter()
        for f in fs:
            orc.set_wall_traction(f)
            ref = orc.apply(C1_RHS, 0.0, tl, cells=False, walls=True)
        t_cpu = (time.perf_counter() - t0) / len(fs)
        out["cpu_baseline"] = {"value": 1.0 / t_cpu, "unit": "wall matvecs/s", "ms_per_matvec": t_cpu * 1e3,
                               "prepare_sing_int_on_wall_ms": t_prep_cpu * 1e3, "cores": os.cpu_count(), "kind": "port",
                               "sample": "the same operator in full on the CPU oracle (OpenMP, all host cores)"}
        out["parity_vs_oracle"] = {"rel_l2_velocity": float(np.linalg.norm(v - ref) / np.linalg.norm(ref))}
    print(json.dumps({"walls": out}))
    return 0


def world_value(v):
    return float(v)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cells", type=int, default=4096)
    ap.add_argument("--seed", type=int, default=161269)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample-cells", type=int, default=32)
    ap.add_argument("--ref-mtube", action="store_true", help="--impl reference: also run the minicase block on the CPU")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-timestep", action="store_true", help="skip the geometry-update / RHS-operator timing")
    ap.add_argument("--host-splines", action="store_true",
                    help="upload spline(g detJ) from the host every step instead of building it on the GPU")
    ap.add_argument("--profile", action="store_true", help="for runs under ncu: exact --warmup, no CPU leg, 1 e2e step")
    ap.add_argument("--no-mtube", action="store_true", help="skip the minicase time-step block (configs[0])")
    ap.add_argument("--mtube-only", action="store_true", help="only the minicase time-step block, one JSON object")
    ap.add_argument("--mtube-steps", type=int, default=4)
    ap.add_argument("--host-noslip", action="store_true", help="mtube block: wall GMRES on the host around per-matvec calls")
    ap.add_argument("--walls-only", action="store_true", help="only the wall-dominated operator block, one JSON object")
    args = ap.parse_args()
    if args.mtube_only:
        return run_mtube(args)
    if args.walls_only:
        return run_walls(args)
    if args.warmup < 3 and args.impl == "b200" and not args.profile:
        args.warmup = 3   # timing hygiene: at least 3 warm-up steps
    if args.profile:
        args.no_cpu_baseline = True
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
