timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "odd_mesh or pme_triple or noncubic" 2>&1 | tail -25 > gpurun_out/t_pytest.log
