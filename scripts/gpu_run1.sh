set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r1_bench_4096.json 2> gpurun_out/r1_bench_4096.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_4096.csv python bench.py --steps 2 --warmup 1 --profile > gpurun_out/r1_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sing_cached|k_pair_self3|k_pair_self|k_spread8|k_interp' -c 8 -o gpurun_out/r1_full_512 -f python bench.py --cells 512 --steps 1 --warmup 1 --profile > gpurun_out/r1_ncu_full.log 2>&1
ls -la gpurun_out
