RBC3D_OVERLAP=1 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/ov_pytest.log
for cells in 512 4096; do
RBC3D_OVERLAP=1 timeout 600 python bench.py --cells $cells --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ov_bench_${cells}.json 2> gpurun_out/ov_bench_${cells}.err
done
