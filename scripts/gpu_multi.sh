# usage: bash scripts/gpu_multi.sh <tag> <ngpu> [cells...]
tag=$1; n=$2; shift; shift
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tests/run_multi_gpu.py > gpurun_out/${tag}_parity.log 2>&1
for cells in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --cells $cells --steps 5 --warmup 3 > gpurun_out/${tag}_bench_${cells}.json 2> gpurun_out/${tag}_bench_${cells}.err
done
