set -x
( nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/r2zb_clocks.csv & echo $! > /tmp/smi.pid )
time python bench.py > gpurun_out/r2zb_bench.json 2> gpurun_out/r2zb_bench.err; echo bench rc=$?
kill $(cat /tmp/smi.pid)
time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2zb_bench_ref.json 2> gpurun_out/r2zb_bench_ref.err; echo ref rc=$?
tail -c 800 gpurun_out/r2zb_bench.err; head -c 2500 gpurun_out/r2zb_bench.json
