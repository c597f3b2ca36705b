python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cta or spacetime or space or staged" 2>&1 | tail -3
python tools/bench_extras.py --only 4 > gpurun_out/r2k_cfg4.jsonl 2> gpurun_out/r2k_cfg4.err; tail -3 gpurun_out/r2k_cfg4.err
python - <<'PY'
import json
for l in open("gpurun_out/r2k_cfg4.jsonl"):
    d = json.loads(l)
    print(d["config"][:110], "| %.3g shots/s" % d["shots_per_s"], "ms", d["ms"], "bp_only", d.get("bp_only", {}).get("ms"), "\n  cta_staged", d.get("bp_only_cta_staged"), "\n  cta_staged_f64", d.get("bp_only_cta_staged_f64"), "\n  f64 sp", d.get("float64"))
PY
