ncu --set full --clock-control none --import-source on -k regex:sweep3d -s 4 -c 1 -o gpurun_out/r1f_sweep3d -f python tools/prof_gibbs3d.py 512 3 > gpurun_out/r1f_ncu3d.log 2>&1
tail -3 gpurun_out/r1f_ncu3d.log
for bz in 48 56 72 96 126; do echo "cfg=6 bz=$bz"; DIM=3 PMG_SW3_BZ=$bz python tools/bench_sweep.py 512 20 2 gibbs; done > gpurun_out/s5_bench.log 2>&1
cat gpurun_out/s5_bench.log
