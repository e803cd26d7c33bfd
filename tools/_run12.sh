python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "fused or 3d or sweep or philox or gibbs or callback or qoi" > gpurun_out/s12_pytest.log 2>&1; tail -3 gpurun_out/s12_pytest.log
DIM=3 python tools/bench_sweep.py 512 20 2 gibbs > gpurun_out/s12_bench.log 2>&1
DIM=3 NOISE=none python tools/bench_sweep.py 512 20 2 gibbs >> gpurun_out/s12_bench.log 2>&1
python tools/bench_sweep.py 4096 40 2 gibbs >> gpurun_out/s12_bench.log 2>&1
cat gpurun_out/s12_bench.log
