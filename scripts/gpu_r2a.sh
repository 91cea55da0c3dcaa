# Round 2, call A (one B200): the whole GPU suite with the new row-walk singular kernel and the reference-config tests,
# bench at 512 and 4096 cells, launch list + one full ncu capture of the singular kernel.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.log
timeout 600 python bench.py --cells 512 --steps 5 --warmup 3 --no-mtube > gpurun_out/r2a_bench_512.json 2> gpurun_out/r2a_bench_512.err
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_4096.json 2> gpurun_out/r2a_bench_4096.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_launches_512.csv \
    python bench.py --cells 512 --steps 2 --warmup 3 --profile --no-mtube > gpurun_out/r2a_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sing_row -s 2 -c 2 -o gpurun_out/r2a_sing_row \
    python bench.py --cells 512 --steps 1 --warmup 3 --profile --no-mtube > gpurun_out/r2a_ncu_full.log 2>&1
ls -la gpurun_out | tail -8
