# Round 2, call Q (two B200s): linear term summed per rank + 3-number all-reduce
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_multi_gpu.py > gpurun_out/r2q_parity.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tests/run_multi_gpu_solver.py > gpurun_out/r2q_solver.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2q_bench_4096_2gpu.json 2> gpurun_out/r2q_bench_4096_2gpu.err
grep -h "multi-gpu\|sharded\|MULTI" gpurun_out/r2q_parity.log gpurun_out/r2q_solver.log
