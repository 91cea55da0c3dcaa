# Round 2, call H (two B200s): sharded device-resident solver, bench e2e through it
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m pytest tests/test_gpu_gmres.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2h_gmres1.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tests/run_multi_gpu_solver.py > gpurun_out/r2h_solver.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_multi_gpu.py > gpurun_out/r2h_parity.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --cells 512 --steps 5 --warmup 3 > gpurun_out/r2h_bench_512.json 2> gpurun_out/r2h_bench_512.err
timeout 900 python bench.py --cells 512 --steps 5 --warmup 3 --no-mtube --no-cpu-baseline > gpurun_out/r2h_bench_512_n1.json 2> gpurun_out/r2h_bench_512_n1.err
cat gpurun_out/r2h_gmres1.log; tail -n 3 gpurun_out/r2h_solver.log; tail -n 2 gpurun_out/r2h_parity.log
