python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2c_pytest.log
cat gpurun_out/r2c_pytest.log
