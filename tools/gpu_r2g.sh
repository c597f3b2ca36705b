python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2g_pytest.log
cat gpurun_out/r2g_pytest.log
F32='{"p":0.05,"shots":8000000,"osd":7,"bp_only":true,"cfg":{"variant":"min_sum","max_iter":100,"alpha":0.8,"damping":0.7,"clip":25.0,"precision":32}}'
F32b='{"code":"[[72, 12, 6]]","p":0.05,"shots":8000000,"osd":0,"bp_only":true,"cfg":{"variant":"min_sum","max_iter":50,"alpha":0.8,"damping":0.7,"clip":25.0,"precision":32}}'
python tools/probe.py "$F32" "$F32b" > gpurun_out/r2g_probe.jsonl 2> gpurun_out/r2g_probe.err; cat gpurun_out/r2g_probe.jsonl | cut -c 200-900
