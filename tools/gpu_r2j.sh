set -x
# ncu captures: float64 warp kernel + OSD on 2M shots (DRAM traffic per shot), the staged CTA kernels on the space-time matrix
F64='{"p":0.05,"shots":2000000,"osd":7,"reps":1,"cfg":{"variant":"min_sum","max_iter":100,"alpha":0.8,"damping":0.7,"clip":25.0,"precision":64}}'
python tools/probe.py "$F64" > gpurun_out/r2j_plain1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'bp_warp_kernel_f64|osd0_fast' -s 2 -c 2 -o gpurun_out/r2j_f64 python tools/probe.py "$F64" > gpurun_out/r2j_ncu1.log 2>&1
python tools/run_stage_probe.py > gpurun_out/r2j_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'bp_stage_kernel|osd0_block_fast' -c 4 -o gpurun_out/r2j_stage python tools/run_stage_probe.py > gpurun_out/r2j_ncu2.log 2>&1
tail -3 gpurun_out/r2j_ncu1.log gpurun_out/r2j_ncu2.log gpurun_out/r2j_plain2.log
