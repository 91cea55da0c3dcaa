# Round 2: PME chain beside the real-space kernels on ONE rank (default for large lists), stage times from a serialised pass
mkdir -p gpurun_out
timeout 500 python bench.py --no-cpu-baseline --no-mtube --steps 10 --warmup 3 > gpurun_out/r2ov_bench.json 2> gpurun_out/r2ov_bench.err
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gmres.py -m gpu -x -q -k "apply or gmres or timings" 2>&1 | tail -2
python - <<'PY'
import json
b=json.loads(open("gpurun_out/r2ov_bench.json").read().strip().splitlines()[-1])
print(b["value"], b["ms_per_step"], b["e2e"]["ms_per_step"], b["critical_paths_ms"])
print(b["stage_ms_mode"]); print({k:round(v,2) for k,v in b["stage_ms"].items() if v>0.2})
print(b["roofline"]); print(b["timestep"]); print(b["parity"])
PY
tail -2 gpurun_out/r2ov_bench.err
