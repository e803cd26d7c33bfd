python -m pytest tests/test_gpu_parity.py -m gpu -q -k "config5" > gpurun_out/s16_pytest.log 2>&1; tail -30 gpurun_out/s16_pytest.log
