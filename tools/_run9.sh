python -m pytest tests -m gpu -x -q > gpurun_out/s10_pytest.log 2>&1; tail -3 gpurun_out/s10_pytest.log
python tools/bench_mg3d.py 513 10 7 > gpurun_out/s10_bench.log 2>&1
python tools/bench_sweep.py 4097 40 8 mg >> gpurun_out/s10_bench.log 2>&1
cat gpurun_out/s10_bench.log
