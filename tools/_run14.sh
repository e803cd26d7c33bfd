python -m pytest tests -m gpu -x -q > gpurun_out/r1h_pytest_gpu.log 2>&1; tail -3 gpurun_out/r1h_pytest_gpu.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r1h_bench.json 2> gpurun_out/r1h_bench.err; tail -c 600 gpurun_out/r1h_bench.json; tail -2 gpurun_out/r1h_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1h_smoke.log 2>&1; tail -2 gpurun_out/r1h_smoke.log
