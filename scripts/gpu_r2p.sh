# Round 2, call P (eight B200s): final multi-GPU numbers -- sharded solver parity at 8 ranks, bench at 8 and at 4 ranks
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 tests/run_multi_gpu_solver.py > gpurun_out/r2p_solver8.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29516 tests/run_multi_gpu_walls.py > gpurun_out/r2p_walls8.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2p_bench_4096_8gpu.json 2> gpurun_out/r2p_bench_4096_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2p_bench_4096_4gpu.json 2> gpurun_out/r2p_bench_4096_4gpu.err
grep -h "sharded\|walls\|MULTI" gpurun_out/r2p_solver8.log gpurun_out/r2p_walls8.log
