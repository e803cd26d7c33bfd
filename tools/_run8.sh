python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gamgmc_3d" > gpurun_out/s9_pytest.log 2>&1; tail -3 gpurun_out/s9_pytest.log
python tools/bench_mg3d.py 513 10 7 > gpurun_out/s9_bench.log 2>&1
cat gpurun_out/s9_bench.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s9_launches_mg3d.csv python tools/bench_mg3d.py 513 2 7 > gpurun_out/s9_ncu.log 2>&1
