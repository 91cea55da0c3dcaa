# Round 2, call O (one B200): cost-proportional CTA assignment in the row kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_reference_configs.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2o_pytest.log
timeout 600 python bench.py --cells 512 --steps 5 --warmup 3 --no-mtube --no-cpu-baseline --no-timestep > gpurun_out/r2o_bench_512.json 2> gpurun_out/r2o_bench_512.err
timeout 900 python bench.py --steps 5 --warmup 3 --no-mtube --no-cpu-baseline > gpurun_out/r2o_bench_4096.json 2> gpurun_out/r2o_bench_4096.err
cat gpurun_out/r2o_pytest.log
