set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_ci.py -m gpu -x -q -k "sum_product or kat or ler" 2>&1 | tail -15
SP='{"p":0.05,"shots":500000,"osd":7,"bp_only":true,"cfg":{"variant":"sum_product","max_iter":100,"precision":64}}'
SPT='{"p":0.05,"shots":100000,"osd":7,"bp_only":true,"cfg":{"variant":"sum_product","max_iter":100,"precision":64,"lanes_per_shot":8}}'
SPS='{"p":0.05,"shots":500000,"osd":7,"bp_only":true,"cfg":{"variant":"sum_product_sym","max_iter":100,"alpha":0.9,"damping":0.8,"clip":20.0,"precision":64}}'
SP72='{"code":"[[72, 12, 6]]","p":0.05,"shots":500000,"osd":0,"bp_only":true,"cfg":{"variant":"sum_product","max_iter":50,"precision":64}}'
SP288='{"code":"[[288, 12, 18]]","p":0.05,"shots":250000,"osd":-1,"cfg":{"variant":"sum_product","max_iter":50,"precision":64}}'
python tools/probe.py "$SP" "$SPT" "$SPS" "$SP72" "$SP288" > gpurun_out/r2n_probe.jsonl 2> gpurun_out/r2n_probe.err
tail -3 gpurun_out/r2n_probe.err
python - <<'PY'
import json
for l in open("gpurun_out/r2n_probe.jsonl"):
    d=json.loads(l); print(d["probe"].get("code","144"), d["probe"]["cfg"], "kernel", d.get("kernel"), "%.3g shots/s %.3g shot-it/s" % (d["shots_per_s"], d["shot_iterations_per_s"]), "bp_only", d.get("bp_only"))
PY
SPN='{"p":0.05,"shots":500000,"osd":-1,"reps":1,"cfg":{"variant":"sum_product","max_iter":100,"precision":64}}'
ncu --set full --clock-control none --import-source on -k regex:'bp_warp_kernel_f64' -s 1 -c 1 -o gpurun_out/r2n_f64sp python tools/probe.py "$SPN" > gpurun_out/r2n_ncu.log 2>&1
tail -2 gpurun_out/r2n_ncu.log
