python tools/bench_csr.py 3 160 20 > gpurun_out/r1h_csr.log 2>&1
python tools/bench_csr.py 3 256 20 >> gpurun_out/r1h_csr.log 2>&1
python tools/bench_csr.py 2 4097 20 >> gpurun_out/r1h_csr.log 2>&1
NOISE=none python tools/bench_csr.py 3 256 20 >> gpurun_out/r1h_csr.log 2>&1
cat gpurun_out/r1h_csr.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sell_sweep -s 6 -c 1 -o gpurun_out/r1h_sell_sweep -f python tools/bench_csr.py 3 256 2 > gpurun_out/r1h_ncu_csr.log 2>&1; tail -2 gpurun_out/r1h_ncu_csr.log
