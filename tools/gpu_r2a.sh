set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.log
F64='{"p":0.05,"shots":2000000,"osd":7,"bp_only":true,"cfg":{"variant":"min_sum","max_iter":100,"alpha":0.8,"damping":0.7,"clip":25.0,"precision":64}}'
F64b='{"code":"[[288, 12, 18]]","p":0.05,"shots":1000000,"osd":-1,"cfg":{"variant":"min_sum","max_iter":50,"alpha":0.8,"damping":0.7,"clip":25.0,"precision":64}}'
F64c='{"code":"[[72, 12, 6]]","p":0.05,"shots":2000000,"osd":0,"cfg":{"variant":"min_sum","max_iter":50,"alpha":0.8,"damping":0.7,"clip":25.0,"precision":64}}'
python tools/probe.py "$F64" "$F64b" "$F64c" > gpurun_out/r2a_probe_words64.jsonl 2> gpurun_out/r2a_probe_words64.err
QLDPC_F64_SPLIT_PLANES=1 python tools/probe.py "$F64" "$F64b" "$F64c" > gpurun_out/r2a_probe_split.jsonl 2> gpurun_out/r2a_probe_split.err
cat gpurun_out/r2a_pytest.log gpurun_out/r2a_probe_split.jsonl gpurun_out/r2a_probe_words64.jsonl
