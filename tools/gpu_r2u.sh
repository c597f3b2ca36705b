set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host_side_packing or pipeline" 2>&1 | tail -5
show() { python - "$1" <<'PY'
import json,sys
raw=open(sys.argv[1]).read(); d=json.loads([l for l in raw.splitlines() if l.startswith("{")][0]); e=d["e2e"]
print(sys.argv[1], "value %.4g e2e %.4g [%s host %d dev %d] forced: device %.4g host %.4g | packed %.4g mc %.4g | f64 %.4g e2e %.4g" % (d["value"], e["value"], e["rows_packed_by"][:12], e["chunks_packed_by_host"], e["chunks_packed_by_device"], e["forced_modes"]["device"]["value"], e["forced_modes"]["host"]["value"], e["packed_host_rows"]["value"], e["mc_sweep"]["value"], d["bit_exact_f64"]["value"], d["bit_exact_f64"]["e2e"]["value"]))
PY
}
python bench.py --no-cpu --no-extras > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; show gpurun_out/r2u_bench.json
QLDPC_HOST_CHUNK=2097152 python bench.py --no-cpu --no-extras > gpurun_out/r2u_bench_2m.json 2>> gpurun_out/r2u_bench.err; show gpurun_out/r2u_bench_2m.json
QLDPC_HOST_CHUNK=524288 python bench.py --no-cpu --no-extras > gpurun_out/r2u_bench_512k.json 2>> gpurun_out/r2u_bench.err; show gpurun_out/r2u_bench_512k.json
tail -3 gpurun_out/r2u_bench.err
