set -x
python bench.py > gpurun_out/r2zn_bench.json 2> gpurun_out/r2zn_bench.err; echo bench rc=$?
python bench.py --impl reference > gpurun_out/r2zn_bench_ref.json 2> gpurun_out/r2zn_bench_ref.err; echo ref rc=$?
python tools/bench_extras.py > gpurun_out/r2zn_extras.jsonl 2> gpurun_out/r2zn_extras.err; echo extras rc=$?
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2zn_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2zn_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2zn_ncu2.log 2>&1
head -c 250 gpurun_out/r2zn_bench.json
