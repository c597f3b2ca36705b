set -x
( nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/r2zi_clocks.csv & echo $! > /tmp/smi.pid )
/usr/bin/time -v python bench.py > gpurun_out/r2zi_bench.json 2> gpurun_out/r2zi_bench.err; echo bench rc=$?
kill $(cat /tmp/smi.pid)
python bench.py --impl reference > gpurun_out/r2zi_bench_ref.json 2> gpurun_out/r2zi_bench_ref.err; echo ref rc=$?
grep -E "Elapsed|Maximum resident" gpurun_out/r2zi_bench.err; head -c 300 gpurun_out/r2zi_bench.json; echo; head -c 300 gpurun_out/r2zi_bench_ref.json
