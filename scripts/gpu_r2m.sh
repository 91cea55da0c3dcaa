# Round 2, call M (one B200): division-free cache build, interpolation kernel at 4 CTAs/SM
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2m_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-mtube --no-cpu-baseline > gpurun_out/r2m_bench_4096.json 2> gpurun_out/r2m_bench_4096.err
cat gpurun_out/r2m_pytest.log
