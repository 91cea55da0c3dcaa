for r in 1 3 5 7 2; do
RBC3D_SING_PITCH_MOD=$r timeout 300 python bench.py --cells 512 --steps 5 --warmup 3 --no-cpu-baseline --no-timestep > gpurun_out/pm${r}_bench_512.json 2> gpurun_out/pm${r}_bench_512.err
done
RBC3D_SING_PITCH_MOD=5 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "singular or apply_matches" 2>&1 | tail -3 > gpurun_out/pm_pytest.log
