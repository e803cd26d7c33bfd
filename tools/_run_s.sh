python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tail or box3" > gpurun_out/r2s_tests.log 2>&1; tail -3 gpurun_out/r2s_tests.log
show() { python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['mgmc_ms_per_sample'], d['launches_per_sample'], [(k['kernel'][:14],round(k['us_per_launch'],1)) for k in d['kernels'] if 'tail' in k['kernel'] or 'Chol' in k['kernel']])"; }
python bench.py --no-cpu-baseline --no-csr --no-gibbs3d --no-mgmc3d --steps 10 2>/dev/null | show
PMG_TAIL_MAX=1500 python bench.py --no-cpu-baseline --no-csr --no-gibbs3d --no-mgmc3d --steps 10 2>/dev/null | show
PMG_TAIL_MAX=400 python bench.py --no-cpu-baseline --no-csr --no-gibbs3d --no-mgmc3d --steps 10 2>/dev/null | show
python bench.py --coarsest-max 1200 --no-cpu-baseline --no-csr --no-gibbs3d --no-mgmc3d --steps 10 2>/dev/null | show
PMG_BOX3_NT=1024 python tools/bench_mg3d.py 513 10 7 2>&1 | head -1
