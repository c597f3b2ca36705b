set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py > gpurun_out/r2zl_bench.json 2> gpurun_out/r2zl_bench.err; echo bench rc=$?
python bench.py --impl reference > gpurun_out/r2zl_bench_ref.json 2> gpurun_out/r2zl_bench_ref.err; echo ref rc=$?
python tools/bench_extras.py > gpurun_out/r2zl_extras.jsonl 2> gpurun_out/r2zl_extras.err; echo extras rc=$?
timeout 120 python tools/osd_block_probe.py > gpurun_out/r2zl_probe.jsonl 2> gpurun_out/r2zl_probe.err && timeout 400 ncu --set full --clock-control none --import-source on -k regex:osd0_block_fast -c 1 -o gpurun_out/r2zl_osdblock python tools/osd_block_probe.py > gpurun_out/r2zl_ncu.log 2>&1
cat gpurun_out/r2zl_probe.jsonl | cut -c1-220
