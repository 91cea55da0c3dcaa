# First GPU call of round 2 (one B200): everything that was added after round 1's GPU budget was spent and has
# therefore only run on the CPU side.  usage: gpurun --timeout 1500 -- 'bash scripts/gpu_r2_first.sh'
mkdir -p gpurun_out
# 1. the GPU tests that have not yet run on a device, one by one so that a failure does not hide the others
for t in "tests/test_gpu_walls.py::test_noslip_wall_solve" \
         "tests/test_gpu_walls.py::test_mtube_time_step" \
         "tests/test_gpu_walls.py::test_case_and_case_sickles_configurations" \
         "tests/test_gpu_walls.py::test_wall_dominated_operator_at_carotid_size" \
         "tests/test_gpu_walls.py::test_device_glob_sph_trans_reproduces_the_reference_exported_cell" \
         "tests/test_gpu_walls.py::test_edge_cases_empty_inactive_and_zero_inputs"; do
  echo "=== $t" >> gpurun_out/r2a_newtests.log
  timeout 600 python -m pytest "$t" -q -m gpu 2>&1 | tail -25 >> gpurun_out/r2a_newtests.log
done
# 2. the whole GPU suite
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2a_pytest.log
# 3+4. the mtube block alone (child process of the bench), then the bench with parity_full_size and mtube in its line
timeout 600 python bench.py --mtube-only > gpurun_out/r2a_mtube.json 2> gpurun_out/r2a_mtube.err
timeout 600 python bench.py --walls-only > gpurun_out/r2a_walls.json 2> gpurun_out/r2a_walls.err
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_4096.json 2> gpurun_out/r2a_bench_4096.err
# 5. (with gpurun --gpus 2) the wall paths over two ranks:
#    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
#        tests/run_multi_gpu_walls.py > gpurun_out/r2a_multi_walls.log 2>&1
