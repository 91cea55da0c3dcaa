# usage: bash scripts/gpu_ncu_k.sh <tag> <kernel-regex> <cells> [count]
tag=$1; rx=$2; cells=$3; cnt=${4:-2}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$rx" -c $cnt -o gpurun_out/${tag} -f python bench.py --cells $cells --steps 1 --warmup 1 --profile > gpurun_out/${tag}.log 2>&1
