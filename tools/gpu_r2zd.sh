set -x
timeout 120 python tools/osd_block_probe.py > gpurun_out/r2zd_probe.jsonl 2> gpurun_out/r2zd_probe.err && timeout 400 ncu --set full --clock-control none --import-source on -k regex:osd0_block_fast -c 1 -o gpurun_out/r2zd_osdblock python tools/osd_block_probe.py > gpurun_out/r2zd_ncu.log 2>&1
cat gpurun_out/r2zd_probe.jsonl
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2zd_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2zd_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2zd_ncu2.log 2>&1
tail -2 gpurun_out/r2zd_ncu2.log | cut -c1-300
