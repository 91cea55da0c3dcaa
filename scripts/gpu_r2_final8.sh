# Round 2, final 8-GPU evidence: parity of the cell operator and the sharded solver at 8 ranks, then the bench line
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/run_multi_gpu.py > gpurun_out/r2f8_parity.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 tests/run_multi_gpu_solver.py > gpurun_out/r2f8_solver.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-mtube > gpurun_out/r2f8_bench_4096.json 2> gpurun_out/r2f8_bench_4096.err
tail -n 3 gpurun_out/r2f8_parity.log gpurun_out/r2f8_solver.log
python - <<'PY'
import json
b=json.loads(open("gpurun_out/r2f8_bench_4096.json").read().strip().splitlines()[-1])
print(b["value"], b["ms_per_step"], b["e2e"]["value"], b.get("critical_paths_ms"), b["stage_ms"])
PY
