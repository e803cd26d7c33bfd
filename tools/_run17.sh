for t in 1 2 4 8; do echo "trep=$t"; PMG_BOX_STREAM_TREP=$t python tools/bench_sweep.py 4097 40 8 mg; done > gpurun_out/s17_bench.log 2>&1
cat gpurun_out/s17_bench.log
