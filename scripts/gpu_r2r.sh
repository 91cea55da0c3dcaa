# Round 2, call R (one B200): row kernel on other mesh sizes, whole GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "other_meshes" 2>&1 | tail -15 > gpurun_out/r2r_meshes.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2r_pytest.log
cat gpurun_out/r2r_meshes.log gpurun_out/r2r_pytest.log
