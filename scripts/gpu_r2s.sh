# Round 2, call S (one B200): wall no-slip GMRES with the matvec replayed from a CUDA graph
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_walls.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2s_walls.log
RBC3D_NOSLIP_GRAPH=0 timeout 600 python bench.py --mtube-only --no-cpu-baseline > gpurun_out/r2s_mtube_eager.json 2> gpurun_out/r2s_mtube_eager.err
timeout 600 python bench.py --mtube-only --no-cpu-baseline > gpurun_out/r2s_mtube_graph.json 2> gpurun_out/r2s_mtube_graph.err
timeout 600 python -m pytest tests/test_gpu_reference_configs.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2s_configs.log
cat gpurun_out/r2s_walls.log gpurun_out/r2s_configs.log
tail -c 1500 gpurun_out/r2s_mtube_eager.json; echo; tail -c 1500 gpurun_out/r2s_mtube_graph.json
tail -3 gpurun_out/r2s_mtube_graph.err
