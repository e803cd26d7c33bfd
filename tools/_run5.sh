python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_3d or 3d or philox" > gpurun_out/s6_pytest.log 2>&1; tail -3 gpurun_out/s6_pytest.log
timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "sweep" >> gpurun_out/s6_pytest.log 2>&1; tail -3 gpurun_out/s6_pytest.log
for bz in 64 84; do echo "cfg=6 bz=$bz"; DIM=3 PMG_SW3_BZ=$bz timeout 120 python tools/bench_sweep.py 512 20 2 gibbs; done > gpurun_out/s6_bench.log 2>&1
cat gpurun_out/s6_bench.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sweep3d -s 1 -c 1 -o gpurun_out/r1f_sweep3d -f python tools/prof_gibbs3d.py 512 3 > gpurun_out/r1f_ncu3d.log 2>&1
tail -3 gpurun_out/r1f_ncu3d.log
