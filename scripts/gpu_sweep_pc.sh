for nt in 384 512; do for pc in 2 3; do
  RBC3D_SING_NT=$nt RBC3D_SING_PC=$pc timeout 600 python bench.py --cells 512 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/nt${nt}_pc${pc}_bench_512.json 2> gpurun_out/nt${nt}_pc${pc}_bench_512.err
done; done
RBC3D_SING_NT=384 RBC3D_SING_PC=2 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "singular or full or operator" 2>&1 | tail -3 > gpurun_out/nt384_pytest.log
