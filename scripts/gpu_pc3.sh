RBC3D_SING_PC=3 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "singular or apply_matches" 2>&1 | tail -4 > gpurun_out/pc3s_pytest.log
for cells in 512 4096; do
RBC3D_SING_PC=3 timeout 600 python bench.py --cells $cells --steps 5 --warmup 3 --no-cpu-baseline --no-timestep > gpurun_out/pc3s_bench_${cells}.json 2> gpurun_out/pc3s_bench_${cells}.err
done
