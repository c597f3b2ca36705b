python -m pytest tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2d_pytest.log
cat gpurun_out/r2d_pytest.log
python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r2d_bench.err; cut -c1-600 gpurun_out/r2d_bench.json
