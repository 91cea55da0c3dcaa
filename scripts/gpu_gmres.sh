timeout 900 python -m pytest tests/test_gpu_gmres.py -x -q -s 2>&1 | tail -25 > gpurun_out/gmres_pytest.log
timeout 600 python bench.py --cells 512 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/gm_bench_512.json 2> gpurun_out/gm_bench_512.err
