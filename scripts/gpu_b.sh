tag=$1; shift
for cells in "$@"; do
  timeout 900 python bench.py --cells $cells --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_${cells}.json 2> gpurun_out/${tag}_bench_${cells}.err
done
