timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "geometry_splines or device_splines" 2>&1 | tail -8 > gpurun_out/g_pytest.log
timeout 900 python bench.py --cells 4096 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench_4096.json 2> gpurun_out/g_bench_4096.err
