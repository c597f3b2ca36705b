set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host_side_packing or pipeline" 2>&1 | tail -5
python bench.py --no-cpu > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo rc=$?
tail -5 gpurun_out/r2r_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2r_bench.json"))
print("value", d["value"], "e2e", json.dumps(d["e2e"], indent=0)[:1800])
print("f64", d["bit_exact_f64"]["value"], json.dumps(d["bit_exact_f64"]["e2e"])[:900])
PY
QLDPC_TRACE=1 python bench.py --no-cpu --steps 3 --warmup 3 2>&1 | grep "qldpc trace" | head -60 > gpurun_out/r2r_trace.txt
