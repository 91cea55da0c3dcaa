timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > gpurun_out/pw_pytest.log
for pc in 2 3; do
  RBC3D_SING_PC=$pc timeout 600 python bench.py --cells 512 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/pw_pc${pc}_bench_512.json 2> gpurun_out/pw_pc${pc}_bench_512.err
done
