# Round 2, call I (eight B200s): slab PME + sharded solver at 8 ranks -- parity, then the headline bench
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/run_multi_gpu.py > gpurun_out/r2i_parity.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 tests/run_multi_gpu_solver.py > gpurun_out/r2i_solver.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench_4096.json 2> gpurun_out/r2i_bench_4096.err
tail -n 3 gpurun_out/r2i_parity.log gpurun_out/r2i_solver.log
