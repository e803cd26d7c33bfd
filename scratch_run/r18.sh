python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "box_colour_pair or gamgmc_3d or coarse_tail" > gpurun_out/s18_pytest.log 2>&1; tail -15 gpurun_out/s18_pytest.log
python tools/bench_mg3d.py 513 10 7 > gpurun_out/s18_bench.log 2>&1
PMG_NO_BOX_PAIR=1 python tools/bench_mg3d.py 513 10 7 >> gpurun_out/s18_bench.log 2>&1
cat gpurun_out/s18_bench.log
