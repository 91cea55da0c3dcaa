# compute-sanitizer over this round's new kernels: row-walk singular kernel (TMA band ring, mbarriers, named barriers),
# slab PME pack / transpose kernels (forced on one rank), closest-neighbour queries, device no-slip solve
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_walls.py -x -q -m gpu \
   -k "singular or apply_matches or slab_decomposed or closest or device_resident_noslip or golden" 2>&1 | tail -25 > gpurun_out/r2_san_memcheck.log
RBC3D_PME_SLAB=1 timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pme_triple or odd_mesh" 2>&1 | tail -25 > gpurun_out/r2_san_memcheck_slab.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_singular" 2>&1 | tail -25 > gpurun_out/r2_san_racecheck.log
tail -n 4 gpurun_out/r2_san_memcheck.log gpurun_out/r2_san_memcheck_slab.log gpurun_out/r2_san_racecheck.log
