timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/fin_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin_smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/fin_bench_4096.json 2> gpurun_out/fin_bench_4096.err
