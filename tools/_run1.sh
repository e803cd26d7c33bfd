set -x
for c in 2 6; do echo "cfg=$c"; DIM=3 PMG_SW3_CFG=$c python tools/bench_sweep.py 512 20 2 gibbs; done > gpurun_out/s2_ws.log 2>&1
for bz in 32 128; do echo "cfg=6 bz=$bz"; DIM=3 PMG_SW3_CFG=6 PMG_SW3_BZ=$bz python tools/bench_sweep.py 512 20 2 gibbs; done >> gpurun_out/s2_ws.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest_gpu.log 2>&1
tail -3 gpurun_out/s2_pytest_gpu.log
cat gpurun_out/s2_ws.log
