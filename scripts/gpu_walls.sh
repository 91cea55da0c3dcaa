set -x
timeout 900 python -m pytest tests/test_gpu_walls.py -x -q 2>&1 | tail -40 > gpurun_out/walls_pytest.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_walls.py -x -q -k "two_walls or wall_matrix" 2>&1 | tail -30 > gpurun_out/walls_sanitizer.log
