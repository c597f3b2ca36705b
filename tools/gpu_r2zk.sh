set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2zk_pytest.log; cat gpurun_out/r2zk_pytest.log
python bench.py > gpurun_out/r2zk_bench.json 2> gpurun_out/r2zk_bench.err; echo bench rc=$?
python tools/bench_extras.py > gpurun_out/r2zk_extras.jsonl 2> gpurun_out/r2zk_extras.err; echo extras rc=$?
grep '"config": "4' gpurun_out/r2zk_extras.jsonl | cut -c1-200
