# Round 2, call T (one B200): cached-graph no-slip solve; per-kernel durations of minicase steps (ncu launch list, eager)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_walls.py tests/test_gpu_reference_configs.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2t_tests.log
timeout 600 python bench.py --mtube-only --no-cpu-baseline --mtube-steps 8 > gpurun_out/r2t_mtube_graph.json 2> gpurun_out/r2t_mtube_graph.err
RBC3D_NOSLIP_GRAPH=0 timeout 900 ncu --target-processes all --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv \
  --log-file gpurun_out/r2t_mtube_launches.csv python bench.py --mtube-only --mtube-steps 1 --no-cpu-baseline > gpurun_out/r2t.log 2>&1
cat gpurun_out/r2t_tests.log
tail -c 900 gpurun_out/r2t_mtube_graph.json
wc -l gpurun_out/r2t_mtube_launches.csv
