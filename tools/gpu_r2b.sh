set -x
F64='{"p":0.05,"shots":500000,"osd":7,"reps":1,"cfg":{"variant":"min_sum","max_iter":100,"alpha":0.8,"damping":0.7,"clip":25.0,"precision":64}}'
python tools/probe.py "$F64" > gpurun_out/r2b_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'bp_warp_kernel_f64|osd0_fast' -s 2 -c 2 -o gpurun_out/r2b_f64 python tools/probe.py "$F64" > gpurun_out/r2b_ncu.log 2>&1
tail -5 gpurun_out/r2b_ncu.log
