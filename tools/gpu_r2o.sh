set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_ci.py tests/test_gpu_experiments.py -m gpu -x -q -k "sum_product or kat or ler or cta_staged or spacetime or per_iteration" 2>&1 | tail -8
SP='{"p":0.05,"shots":500000,"osd":7,"bp_only":true,"cfg":{"variant":"sum_product","max_iter":100,"precision":64}}'
SPS='{"p":0.05,"shots":500000,"osd":7,"bp_only":true,"cfg":{"variant":"sum_product_sym","max_iter":100,"alpha":0.9,"damping":0.8,"clip":20.0,"precision":64}}'
python tools/probe.py "$SP" "$SPS" > gpurun_out/r2o_probe.jsonl 2> gpurun_out/r2o_probe.err
tail -3 gpurun_out/r2o_probe.err
python tools/bench_extras.py --only 4 > gpurun_out/r2o_cfg4.jsonl 2> gpurun_out/r2o_cfg4.err; tail -3 gpurun_out/r2o_cfg4.err
python - <<'PY'
import json
for l in open("gpurun_out/r2o_probe.jsonl"):
    d=json.loads(l); print(d["probe"].get("code","144"), d["probe"]["cfg"]["variant"], "kernel", d.get("kernel"), "%.3g shots/s %.3g shot-it/s" % (d["shots_per_s"], d["shot_iterations_per_s"]), "bp_only", d.get("bp_only"))
for l in open("gpurun_out/r2o_cfg4.jsonl"):
    d = json.loads(l)
    print(d["config"][:110], "| %.3g shots/s" % d["shots_per_s"], "ms", d["ms"], "bp_only", d.get("bp_only", {}).get("ms"), "\n  f64 sp", d.get("float64"))
PY
