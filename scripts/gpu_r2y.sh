# Round 2, call Y (two B200): multi-GPU parity, walls and sharded solver after the GMRES / spreading changes; 2-GPU bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "several_gpus" 2>&1 | tail -6 > gpurun_out/r2y_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-mtube > gpurun_out/r2y_bench2.json 2> gpurun_out/r2y_bench2.err
cat gpurun_out/r2y_multi.log
python - <<'PY'
import json
try:
    b=json.loads(open("gpurun_out/r2y_bench2.json").read().strip().splitlines()[-1])
    print("bench2", b["value"], b["ms_per_step"], b["e2e"]["value"], b.get("parity"))
except Exception as e: print("bench err",e)
PY
tail -3 gpurun_out/r2y_bench2.err
