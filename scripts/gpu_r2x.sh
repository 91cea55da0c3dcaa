# Round 2, call X (one B200): run-ahead GMRES of the wall solve
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_walls.py tests/test_gpu_reference_configs.py tests/test_gpu_gmres.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2x_tests.log
RBC3D_NOSLIP_RUN_AHEAD=0 timeout 600 python bench.py --mtube-only --no-cpu-baseline --mtube-steps 8 > gpurun_out/r2x_mtube_sync.json 2> gpurun_out/r2x_mtube_sync.err
timeout 600 python bench.py --mtube-only --no-cpu-baseline --mtube-steps 8 > gpurun_out/r2x_mtube.json 2> gpurun_out/r2x_mtube.err
RBC3D_NOSLIP_GRAPH=0 RBC3D_NOSLIP_RUN_AHEAD=0 timeout 900 ncu --target-processes all --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv \
  --log-file gpurun_out/r2x_mtube_launches.csv python bench.py --mtube-only --mtube-steps 1 --no-cpu-baseline > gpurun_out/r2x.log 2>&1
cat gpurun_out/r2x_tests.log
python - <<'PY'
import json
for f in ("r2x_mtube_sync","r2x_mtube"):
    d=json.load(open(f"gpurun_out/{f}.json"))["mtube"]
    print(f, d["bi_timesteps_per_s"], [[round(x,2) for x in r] for r in d["ms_geometry_rhs_noslip"]][-3:], d["wall_gmres_iterations"][-3:])
PY
