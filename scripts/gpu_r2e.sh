# Round 2, call E: row kernel v2 (producer warp, no CTA barriers, L2 prefetch) -- parity, bench, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_gmres.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2e_pytest.log
timeout 900 python -m pytest tests/test_gpu_reference_configs.py -m gpu -x -q -s 2>&1 | tail -25 > gpurun_out/r2e_pytest_ref.log
timeout 600 python bench.py --cells 512 --steps 5 --warmup 3 --no-mtube --no-cpu-baseline > gpurun_out/r2e_bench_512.json 2> gpurun_out/r2e_bench_512.err
timeout 900 python bench.py --steps 5 --warmup 3 --no-mtube --no-cpu-baseline > gpurun_out/r2e_bench_4096.json 2> gpurun_out/r2e_bench_4096.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sing_row -s 2 -c 1 -o gpurun_out/r2e_sing_row \
    python bench.py --cells 512 --steps 1 --warmup 3 --profile --no-mtube > gpurun_out/r2e_ncu_full.log 2>&1
ls -la gpurun_out | tail -6
