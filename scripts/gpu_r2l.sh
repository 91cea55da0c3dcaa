# Round 2, call L (one B200): row kernel with predicated node loads
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "singular or apply_matches or golden" 2>&1 | tail -5 > gpurun_out/r2l_pytest.log
timeout 600 python bench.py --cells 512 --steps 5 --warmup 3 --no-mtube --no-cpu-baseline --no-timestep > gpurun_out/r2l_bench_512.json 2> gpurun_out/r2l_bench_512.err
timeout 900 python bench.py --steps 5 --warmup 3 --no-mtube --no-cpu-baseline --no-timestep > gpurun_out/r2l_bench_4096.json 2> gpurun_out/r2l_bench_4096.err
cat gpurun_out/r2l_pytest.log
