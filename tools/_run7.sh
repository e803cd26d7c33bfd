ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s8_launches_mg3d.csv python tools/bench_mg3d.py 513 2 7 > gpurun_out/s8_ncu.log 2>&1
tail -2 gpurun_out/s8_ncu.log
