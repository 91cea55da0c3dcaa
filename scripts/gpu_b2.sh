timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r1f_bench_4096.json 2> gpurun_out/r1f_bench_4096.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1f_bench_ref.json 2> gpurun_out/r1f_bench_ref.err
