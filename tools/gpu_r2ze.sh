set -x
python bench.py > gpurun_out/r2ze_bench.json 2> gpurun_out/r2ze_bench.err; echo bench rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2ze_bench_ref.json 2> gpurun_out/r2ze_bench_ref.err; echo ref rc=$?
python tools/bench_extras.py > gpurun_out/r2ze_extras.jsonl 2> gpurun_out/r2ze_extras.err; echo extras rc=$?
tail -c 600 gpurun_out/r2ze_bench.err; head -c 600 gpurun_out/r2ze_bench.json; wc -l gpurun_out/r2ze_extras.jsonl
