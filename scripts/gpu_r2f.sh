# Round 2, call F (one B200): slab-decomposed PME forced on one rank + the whole GPU suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "slab_decomposed" 2>&1 | tail -15 > gpurun_out/r2f_slab.log
RBC3D_PME_SLAB=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_walls.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2f_pytest_slabforced.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2f_pytest.log
RBC3D_PME_SLAB=1 timeout 600 python bench.py --cells 512 --steps 5 --warmup 3 --no-mtube --no-cpu-baseline > gpurun_out/r2f_bench_512_slab.json 2> gpurun_out/r2f_bench_512_slab.err
cat gpurun_out/r2f_slab.log
