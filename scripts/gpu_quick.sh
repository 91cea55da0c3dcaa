# usage: bash scripts/gpu_quick.sh <tag> [cells...]  -- all GPU parity tests + bench without the CPU leg
tag=$1; shift
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_pytest.log
for cells in "$@"; do
  timeout 600 python bench.py --cells $cells --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_${cells}.json 2> gpurun_out/${tag}_bench_${cells}.err
done
