# compute-sanitizer over the kernels added after gpu_r2_sanitize.sh: split same-surface pair kernel (few cells),
# source-block spreading chosen per list, one-warp-per-target interpolation, single-launch dot products with tickets,
# the graph-replayed / run-ahead wall solve, re-meshing a live context
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_walls.py tests/test_gpu_gmres.py -x -q -m gpu \
   -k "pair_sum or pme_triple or odd_mesh or raw_targets or noslip or pencil_walk or second_mesh or gmres or mtube_time_step" 2>&1 | tail -25 > gpurun_out/r2_san2_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_walls.py -x -q -m gpu -k "device_resident_noslip or pencil_walk" 2>&1 | tail -25 > gpurun_out/r2_san2_racecheck.log
tail -n 4 gpurun_out/r2_san2_memcheck.log gpurun_out/r2_san2_racecheck.log
