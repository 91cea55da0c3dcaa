# sweep the run-time knobs of the cached singular kernel at 512 cells
for pc in 2 3 4; do for nt in 256 384 512; do
  RBC3D_SING_PC=$pc RBC3D_SING_NT=$nt timeout 300 python bench.py --cells 512 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ss_pc${pc}_nt${nt}.json 2> gpurun_out/ss_pc${pc}_nt${nt}.err
done; done
RBC3D_SING_PER_WARP=1 timeout 300 python bench.py --cells 512 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ss_pw.json 2> gpurun_out/ss_pw.err
