python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused_3d or 3d or philox" > gpurun_out/s4_pytest.log 2>&1; tail -3 gpurun_out/s4_pytest.log
python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "sweep" >> gpurun_out/s4_pytest.log 2>&1; tail -3 gpurun_out/s4_pytest.log
for c in 2 6; do echo "cfg=$c"; DIM=3 PMG_SW3_CFG=$c python tools/bench_sweep.py 512 20 2 gibbs; done > gpurun_out/s4_bench.log 2>&1
echo "cfg=6 bz=84"; DIM=3 PMG_SW3_CFG=6 PMG_SW3_BZ=84 python tools/bench_sweep.py 512 20 2 gibbs >> gpurun_out/s4_bench.log 2>&1
echo "cfg=6 none"; NOISE=none DIM=3 PMG_SW3_CFG=6 python tools/bench_sweep.py 512 20 2 gibbs >> gpurun_out/s4_bench.log 2>&1
cat gpurun_out/s4_bench.log
