# round-1 evidence run, v6: GPU parity, smoke, bench at 4096 cells (with CPU leg), reference arm, launch list, ncu --set full
set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1e_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1e_smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r1e_bench_4096.json 2> gpurun_out/r1e_bench_4096.err
timeout 600 python bench.py --cells 512 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1e_bench_512.json 2> gpurun_out/r1e_bench_512.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r1e_bench_ref.json 2> gpurun_out/r1e_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1e_launches_4096.csv python bench.py --steps 2 --warmup 1 --profile > gpurun_out/r1e_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sing_band|k_pair_self_cached|k_spread_walk|k_interp_walk|k_pair_list|k_spline_build' -c 8 -o gpurun_out/r1e_full_512 -f python bench.py --cells 512 --steps 1 --warmup 1 --profile > gpurun_out/r1e_ncu_full.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_sing_band|k_pair_self_cached' -c 2 --csv --log-file gpurun_out/r1e_traffic_4096.csv python bench.py --steps 1 --warmup 1 --profile > gpurun_out/r1e_ncu_traffic.log 2>&1
ls -la gpurun_out
