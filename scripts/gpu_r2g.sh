# Round 2, call G (two B200s): slab-decomposed PME over NCCL -- parity of every entry point, wall paths, bench at 2 ranks
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_multi_gpu.py > gpurun_out/r2g_parity.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/run_multi_gpu_walls.py > gpurun_out/r2g_parity_walls.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --cells 512 --steps 5 --warmup 3 > gpurun_out/r2g_bench_512.json 2> gpurun_out/r2g_bench_512.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2g_bench_4096.json 2> gpurun_out/r2g_bench_4096.err
tail -3 gpurun_out/r2g_parity.log gpurun_out/r2g_parity_walls.log
