for lr in 2; do
RBC3D_SING_LR=$lr timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5 > gpurun_out/lr${lr}_pytest.log
for cells in 512 4096; do
RBC3D_SING_LR=$lr timeout 600 python bench.py --cells $cells --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/lr${lr}_bench_${cells}.json 2> gpurun_out/lr${lr}_bench_${cells}.err
done
done
