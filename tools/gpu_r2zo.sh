set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py > gpurun_out/r2zo_bench.json 2> gpurun_out/r2zo_bench.err; echo bench rc=$?
python bench.py --impl reference > gpurun_out/r2zo_bench_ref.json 2> gpurun_out/r2zo_bench_ref.err; echo ref rc=$?
python tools/bench_extras.py --only 2,5 > gpurun_out/r2zo_extras_25.jsonl 2>/dev/null
python -c "
import json
for l in open('gpurun_out/r2zo_extras_25.jsonl'):
    c=json.loads(l)
    if 'p=0.01' in c['config']: print(c['config'][:75], '%.4g'%c['shots_per_s'])
"
