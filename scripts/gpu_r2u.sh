# Round 2, call U (one B200): minicase step with the graph cache (debug prints) and with the block spreading kernel
mkdir -p gpurun_out
RBC3D_DEBUG_GRAPH=1 timeout 600 python bench.py --mtube-only --no-cpu-baseline --mtube-steps 8 > gpurun_out/r2u_graph.json 2> gpurun_out/r2u_graph.err
RBC3D_SPREAD_BLOCKS=1 timeout 600 python bench.py --mtube-only --no-cpu-baseline --mtube-steps 8 > gpurun_out/r2u_blocks.json 2> gpurun_out/r2u_blocks.err
RBC3D_SPREAD_BLOCKS=1 RBC3D_NOSLIP_GRAPH=0 timeout 900 ncu --target-processes all --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv \
  --log-file gpurun_out/r2u_blocks_launches.csv python bench.py --mtube-only --mtube-steps 1 --no-cpu-baseline > gpurun_out/r2u.log 2>&1
grep -c "noslip graph" gpurun_out/r2u_graph.err; grep "noslip graph" gpurun_out/r2u_graph.err | tail -12
python - <<'PY'
import json
for f in ("r2u_graph","r2u_blocks"):
    d=json.load(open(f"gpurun_out/{f}.json"))["mtube"]
    print(f, d["bi_timesteps_per_s"], [[round(x,1) for x in r] for r in d["ms_geometry_rhs_noslip"]])
PY
