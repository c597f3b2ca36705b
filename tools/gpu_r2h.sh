python -m pytest tests/test_gpu_ci.py tests/test_gpu_experiments.py tests/test_gpu_parity.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -45 > gpurun_out/r2h_pytest.log
cat gpurun_out/r2h_pytest.log
