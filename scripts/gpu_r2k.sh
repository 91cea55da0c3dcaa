# Round 2, call K (one B200): device-resident NoSlipWall, bench side blocks on the reference meshes, overlap on one GPU
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_walls.py -m gpu -x -q -k "noslip" 2>&1 | tail -15 > gpurun_out/r2k_noslip.log
timeout 600 python bench.py --mtube-only > gpurun_out/r2k_mtube.json 2> gpurun_out/r2k_mtube.err
timeout 600 python bench.py --mtube-only --host-noslip --no-cpu-baseline > gpurun_out/r2k_mtube_host.json 2> gpurun_out/r2k_mtube_host.err
timeout 600 python bench.py --walls-only > gpurun_out/r2k_walls.json 2> gpurun_out/r2k_walls.err
RBC3D_OVERLAP=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-mtube --no-cpu-baseline --no-timestep > gpurun_out/r2k_bench_4096_overlap.json 2> gpurun_out/r2k_bench_4096_overlap.err
cat gpurun_out/r2k_noslip.log
