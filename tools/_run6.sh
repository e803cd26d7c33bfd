python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gamgmc or fused_3d or 3d" > gpurun_out/s7_pytest.log 2>&1; tail -3 gpurun_out/s7_pytest.log
python tools/bench_mg3d.py 513 10 7 > gpurun_out/s7_bench.log 2>&1
PMG_NO_FUSED_MG3=1 python tools/bench_mg3d.py 513 10 7 >> gpurun_out/s7_bench.log 2>&1
DIM=3 python tools/bench_sweep.py 512 20 2 gibbs >> gpurun_out/s7_bench.log 2>&1
DIM=3 python tools/bench_sweep.py 513 20 2 gibbs >> gpurun_out/s7_bench.log 2>&1
cat gpurun_out/s7_bench.log
