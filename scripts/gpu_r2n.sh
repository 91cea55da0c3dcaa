# Round 2, call N (one B200): the round's evidence run -- GPU suite, smoke, reference arm (one complete CPU matvec),
# product arm with the all-targets parity check, launch list at 4096 cells, full ncu capture at 512 cells, DRAM traffic of
# the two largest kernels at 4096 cells
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2n_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2n_smoke.log 2>&1
timeout 1500 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2n_bench_reference.json 2> gpurun_out/r2n_bench_reference.err
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2n_bench_4096.json 2> gpurun_out/r2n_bench_4096.err
timeout 600 python bench.py --cells 512 --steps 10 --warmup 3 --no-mtube > gpurun_out/r2n_bench_512.json 2> gpurun_out/r2n_bench_512.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2n_launches_4096.csv \
    python bench.py --steps 2 --warmup 3 --profile --no-mtube --no-timestep > gpurun_out/r2n_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sing_row|k_pair_self_cached|k_spread_walk|k_interp_walk|k_spline_build|k_pair_list' -s 12 -c 8 -o gpurun_out/r2n_full_512 \
    python bench.py --cells 512 --steps 1 --warmup 3 --profile --no-mtube --no-timestep > gpurun_out/r2n_ncu_full.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_sing_row|k_pair_self_cached' -s 4 -c 2 --csv --log-file gpurun_out/r2n_traffic_4096.csv \
    python bench.py --steps 1 --warmup 3 --profile --no-mtube --no-timestep > gpurun_out/r2n_ncu_traffic.log 2>&1
cat gpurun_out/r2n_pytest.log; tail -n 3 gpurun_out/r2n_smoke.log
