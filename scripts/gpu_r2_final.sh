# Round 2, final single-GPU evidence: whole GPU suite, the default bench line (CPU sample, time step, minicase and
# carotid walls blocks), 512-cell line
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2f_pytest.log
timeout 1500 python bench.py > gpurun_out/r2f_bench_4096.json 2> gpurun_out/r2f_bench_4096.err
timeout 600 python bench.py --cells 512 --no-cpu-baseline --no-mtube > gpurun_out/r2f_bench_512.json 2> gpurun_out/r2f_bench_512.err
cat gpurun_out/r2f_pytest.log
tail -c 2500 gpurun_out/r2f_bench_4096.json
