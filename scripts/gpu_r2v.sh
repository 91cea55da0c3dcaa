# Round 2, call V (one B200): small-configuration latency work (split self-pair kernel, lazy pair cache, per-list
# spreading kernel, direct interpolation, single-sync GMRES, graph replay of the wall matvec)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2v_pytest.log
timeout 600 python bench.py --mtube-only --no-cpu-baseline --mtube-steps 8 > gpurun_out/r2v_mtube.json 2> gpurun_out/r2v_mtube.err
timeout 600 python bench.py --walls-only --no-cpu-baseline > gpurun_out/r2v_walls_auto.json 2> gpurun_out/r2v_walls_auto.err
RBC3D_SPREAD_BLOCKS=0 timeout 600 python bench.py --walls-only --no-cpu-baseline > gpurun_out/r2v_walls_walk.json 2> gpurun_out/r2v_walls_walk.err
timeout 900 python bench.py --no-cpu-baseline --no-mtube > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
cat gpurun_out/r2v_pytest.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2v_mtube.json"))["mtube"]
print("mtube", d["bi_timesteps_per_s"], [[round(x,2) for x in r] for r in d["ms_geometry_rhs_noslip"]][-3:], d["wall_gmres_iterations"][-3:])
for f in ("auto","walk"):
    try:
        t=open(f"gpurun_out/r2v_walls_{f}.json").read(); print("walls",f,t[-700:])
    except Exception as e: print(f,e)
try:
    b=json.loads(open("gpurun_out/r2v_bench.json").read().strip().splitlines()[-1])
    print("bench", b["value"], b["ms_per_step"], b["e2e"], b.get("parity"))
    print({k:v for k,v in b.items() if "timestep" in k or "stage" in k})
except Exception as e: print("bench err",e)
PY
tail -3 gpurun_out/r2v_bench.err
