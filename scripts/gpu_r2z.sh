# quick single-GPU check: bench line without the CPU legs; walls block (warm PrepareSingIntOnWall)
mkdir -p gpurun_out
timeout 900 python bench.py --no-cpu-baseline --no-mtube > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
timeout 600 python bench.py --walls-only --no-cpu-baseline > gpurun_out/r2g_walls.json 2> gpurun_out/r2g_walls.err
python - <<'PY'
import json
b=json.loads(open("gpurun_out/r2g_bench.json").read().strip().splitlines()[-1])
print(b["value"], b["ms_per_step"], b["roofline"]["kernel"], b["timestep"]["geometry_update_ms"], b["timestep"]["gmres_solve_ms"])
w=json.load(open("gpurun_out/r2g_walls.json"))["walls"]
print(w["wall_matvecs_per_s"], w["prepare_sing_int_on_wall_ms"], w["set_walls_and_first_prepare_ms"])
PY
