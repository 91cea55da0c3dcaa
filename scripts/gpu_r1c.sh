# full round-1 evidence run: GPU parity, bench at 4096 cells (with CPU leg), launch list, ncu --set full of top kernels
set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1c_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r1c_bench_4096.json 2> gpurun_out/r1c_bench_4096.err
timeout 600 python bench.py --cells 512 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_bench_512.json 2> gpurun_out/r1c_bench_512.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1c_launches_4096.csv python bench.py --steps 2 --warmup 1 --profile > gpurun_out/r1c_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sing|k_pair|k_spread|k_interp' -c 10 -o gpurun_out/r1c_full_512 -f python bench.py --cells 512 --steps 1 --warmup 1 --profile > gpurun_out/r1c_ncu_full.log 2>&1
ls -la gpurun_out
