timeout 900 python -m pytest tests/test_gpu_gmres.py -x -q -s 2>&1 | tail -25 > gpurun_out/gmres_pytest.log
