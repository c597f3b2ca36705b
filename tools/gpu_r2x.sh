set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "osd or spacetime or arbitrary_sparse or large_check" 2>&1 | tail -8
timeout 300 python tools/osd_block_probe.py > gpurun_out/r2x_probe.jsonl 2> gpurun_out/r2x_probe.err; cat gpurun_out/r2x_probe.jsonl; tail -3 gpurun_out/r2x_probe.err
timeout 120 python tools/run_osd_block.py 2960
