set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2m_pytest.log
python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo bench rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2m_bench_ref.json 2> gpurun_out/r2m_bench_ref.err; echo ref rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2m_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2m_ncu.log 2>&1
cat gpurun_out/r2m_pytest.log; tail -c 1500 gpurun_out/r2m_bench.err; head -c 3000 gpurun_out/r2m_bench.json
