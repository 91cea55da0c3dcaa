"""bench.py contract checks that need no GPU: the reference arm (CPU restatement on the host cores) prints one JSON line
with the agreed keys, ranks other than 0 leave without work, and the product arm refuses to run without a device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e, capture_output=True,
                          text=True, timeout=600)


def test_reference_arm_prints_the_contract_line():
    """One complete CPU matvec, measured (nothing extrapolated), steps = the applications actually run, the same config
    keys as the product arm, all host threads even under torchrun's OMP_NUM_THREADS=1, and the result kept for the GPU
    arm's all-targets check."""
    import numpy as np
    r = _run(["--impl", "reference", "--cells", "8", "--steps", "20", "--warmup", "5", "--seed", "7"],
             env={"OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "Ewald BI matvecs/s" and line["unit"] == "matvecs/s"
    assert line["higher_is_better"] is True and line["dtype"] == "f64"
    assert line["steps"] == 1 and line["warmup"] == 0 and line["steps_requested"] == 20 and line["extrapolated"] is False
    assert line["value"] > 0 and abs(line["value"] * line["ms_per_step"] / 1e3 - 1.0) < 1e-9
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"] and "nothing extrapolated" in cb["sample"]
    assert cb["cores"] == len(os.sched_getaffinity(0))
    assert abs(sum(cb["stage_s"].values()) - line["ms_per_step"] / 1e3) < 0.05 * line["ms_per_step"] / 1e3 + 0.05
    assert line["e2e"] == {"value": line["value"], "unit": "matvecs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cfg = line["config"]
    assert set(cfg) == {"workload", "cells", "points", "alpha", "eps", "P", "rc", "Nc", "visc_ratio", "seed"}
    assert cfg["cells"] == 8 and cfg["P"] == 8 and abs(cfg["rc"] - 1.1986122199272495) < 1e-12
    pth = os.path.join(ROOT, "oracle", "_cache", "ref_matvec_c8_s7.npy")
    try:
        v = np.load(pth)
        assert v.shape == (3, cfg["points"]) and np.isfinite(v).all() and np.abs(v).max() > 0
    finally:
        for q in (pth, pth + ".json"):
            if os.path.exists(q):
                os.remove(q)


def test_reference_arm_does_not_map_the_product_library():
    code = ("import sys, argparse; sys.path.insert(0, %r); import importlib.util as u;"
            "sp = u.spec_from_file_location('b', %r); b = u.module_from_spec(sp); sp.loader.exec_module(b);"
            "import io, contextlib; buf = io.StringIO();\n"
            "with contextlib.redirect_stdout(buf):\n"
            "    b.run_reference(argparse.Namespace(cells=8, seed=11, gpus=1, steps=1, warmup=0, ref_mtube=False))\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'librbc3d_oracle' in maps and 'librbc3d_b200' not in maps, 'product library mapped'\n"
            "import os\n"
            "[os.remove(p) for p in (b.ref_cache_path(8, 11), b.ref_cache_path(8, 11) + '.json') if os.path.exists(p)]\n"
            % (ROOT, os.path.join(ROOT, "bench.py")))
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]


def test_reference_arm_other_ranks_exit_without_work():
    r = _run(["--impl", "reference", "--cells", "8", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run(["--cells", "8", "--steps", "1", "--warmup", "0"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_mtube_block_refuses_without_a_device_and_the_parent_survives_it():
    """--mtube-only is the child process of the main bench (BASELINE.json configs[0] time-step block): without a device
    it refuses like the product arm, and the parent turns a failing child into {"error": ...} instead of dying."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run(["--mtube-only"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    import argparse
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    out = bench.mtube_child(argparse.Namespace(seed=1, mtube_steps=2, no_cpu_baseline=True))
    assert set(out) == {"error"} and "no CUDA device" in out["error"]


def test_mtube_block_python_path_with_a_stand_in_library(monkeypatch, capsys, oracle_lib):
    """run_mtube end to end on a machine without a GPU: EwaldOperator replaced by a stand-in that answers the same calls
    from the oracle, so that the block's own logic (steps, warm-up, JSON keys, parity record) is exercised here."""
    import argparse
    import importlib.util
    import numpy as np
    from rbc3d_b200 import ewald
    from rbc3d_b200.capi import TL_CELLS, TL_WALLS

    class StandIn:
        def __init__(self, Lb, device=-1):
            self.orc = oracle_lib.Oracle(Lb)
            self.Nb = self.orc.Nb
            self.n = 0

        def set_suspension(self, sus):
            self.sus = sus
            self.orc.set_cells(sus)

        def SourceList_UpdateCoord(self, *a):
            self.orc.set_cells(self.sus)

        def SourceList_UpdateDensity(self, *a):
            pass

        def set_walls(self, W):
            self.orc.set_walls(W)

        def PrepareSingIntOnWall(self):
            self.orc.prepare_sing_int_on_walls()

        def set_wall_traction(self, f):
            self.orc.set_wall_traction(f)

        def apply(self, c1, c2, tlist, cells=True, walls=False):
            self.n += 1
            tl = self.orc.cell_targets() if tlist == TL_CELLS else self.orc.wall_targets()
            assert tlist in (TL_CELLS, TL_WALLS)
            return self.orc.apply(c1, c2, tl, cells=cells, walls=walls)

        def launch_count(self):
            return 10 * self.n

        def close(self):
            pass

    monkeypatch.setattr(ewald, "EwaldOperator", StandIn)
    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    monkeypatch.setattr(bench, "MTUBE_WARM_STEPS", 2)
    assert bench.run_mtube(argparse.Namespace(seed=161269, mtube_steps=2, no_cpu_baseline=False, host_noslip=True)) == 0
    m = json.loads(capsys.readouterr().out.strip().splitlines()[-1])["mtube"]
    assert m["steps"] == 2 and len(m["ms_per_step"]) == 2 and m["bi_timesteps_per_s"] > 0 and m["gpu_launches"] > 0
    assert len(m["wall_gmres_iterations"]) == 4 and 0 < m["wall_gmres_iterations"][0] <= 60
    assert "new_cyl_D6_L13_33.e" in m["workload"] and "1328 vertices / 2404 triangles" in m["workload"]
    cb = m["cpu_baseline"]
    assert cb["kind"] == "port" and cb["unit"] == "timesteps/s" and cb["wall_gmres_iterations"] == m["wall_gmres_iterations"]
    par = m["parity_vs_oracle"]
    assert all(par["same_iterations"]) and max(par["rel_l2_cell_velocity"]) < 1e-12      # the stand-in IS the oracle


def test_walls_block_python_path_with_a_stand_in_library(monkeypatch, capsys, oracle_lib):
    """run_walls (configs[4], wall-dominated operator) end to end without a GPU: the library replaced by an
    oracle-backed stand-in, so that the block's own logic and JSON keys are exercised here."""
    import argparse
    import importlib.util
    import numpy as np
    from rbc3d_b200 import ewald

    class StandIn:
        def __init__(self, Lb, device=-1):
            self.orc = oracle_lib.Oracle(Lb)
            self.Nb = self.orc.Nb
            self.n = 0

        def set_walls(self, W):
            self.W = W
            self.orc.set_walls(W, ncell=0)

        def PrepareSingIntOnWall(self):
            self.orc.prepare_sing_int_on_walls()

        def set_wall_traction(self, f):
            self.orc.set_wall_traction(f)

        def apply(self, c1, c2, tlist, cells=True, walls=False):
            self.n += 1
            return self.orc.apply(c1, c2, self.orc.wall_targets(), cells=cells, walls=walls)

        def wall_matrix(self):
            n = sum(len(self.orc.wall_matrix(w)[1]) for w in range(self.W.nwall))
            return np.array([0, n]), None, None

        def launch_count(self):
            return 7 * self.n

        def close(self):
            pass

    monkeypatch.setattr(ewald, "EwaldOperator", StandIn)
    spec = importlib.util.spec_from_file_location("bench_mod3", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.run_walls(argparse.Namespace(seed=161269, no_cpu_baseline=False)) == 0
    w = json.loads(capsys.readouterr().out.strip().splitlines()[-1])["walls"]
    assert w["wall_matvecs_per_s"] > 0 and w["steps"] == 20 and w["gpu_launches"] == 140
    assert w["matrix_blocks_3x3"] > 1_000_000 and "carotid.e + web.e" in w["workload"] and "14550 + 2903" in w["workload"]
    assert w["cpu_baseline"]["kind"] == "port" and w["cpu_baseline"]["value"] > 0
    assert w["parity_vs_oracle"]["rel_l2_velocity"] < 1e-12              # the stand-in IS the oracle
