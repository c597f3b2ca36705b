python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2f_pytest.log
cat gpurun_out/r2f_pytest.log
F64='{"p":0.05,"shots":4000000,"osd":7,"bp_only":true,"cfg":{"variant":"min_sum","max_iter":100,"alpha":0.8,"damping":0.7,"clip":25.0,"precision":64}}'
F64b='{"code":"[[288, 12, 18]]","p":0.05,"shots":1000000,"osd":0,"bp_only":true,"cfg":{"variant":"min_sum","max_iter":50,"alpha":0.8,"damping":0.7,"clip":25.0,"precision":64}}'
python tools/probe.py "$F64" "$F64b" > gpurun_out/r2f_probe.jsonl 2> gpurun_out/r2f_probe.err; cat gpurun_out/r2f_probe.jsonl | cut -c 200-900
