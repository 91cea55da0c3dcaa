# Round 2, call J (one B200): closest-neighbour kernels (8(f)-4) + whole GPU suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_walls.py -m gpu -x -q -k closest 2>&1 | tail -15 > gpurun_out/r2j_closest.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2j_pytest.log
cat gpurun_out/r2j_closest.log gpurun_out/r2j_pytest.log
