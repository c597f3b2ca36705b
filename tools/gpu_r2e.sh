python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2e_pytest.log
cat gpurun_out/r2e_pytest.log
python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r2e_bench.err; cut -c1-300 gpurun_out/r2e_bench.json
