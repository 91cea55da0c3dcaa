# Round 2, call W (one B200): spreading-kernel crossover
mkdir -p gpurun_out
timeout 900 python scripts/spread_crossover.py > gpurun_out/r2w_spread_crossover.jsonl 2> gpurun_out/r2w.err
cat gpurun_out/r2w_spread_crossover.jsonl; tail -3 gpurun_out/r2w.err
