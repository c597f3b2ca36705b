N=${1:-2}
set -x
python tools/host_probe.py > gpurun_out/r2t_host_probe_${N}gpu.json 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR tools/pcie_bw_multi.py > gpurun_out/r2t_pcie_${N}gpu.jsonl 2> gpurun_out/r2t_pcie_${N}gpu.err
QLDPC_PIN_LOCAL=1 $TR tools/pcie_bw_multi.py >> gpurun_out/r2t_pcie_${N}gpu.jsonl 2>> gpurun_out/r2t_pcie_${N}gpu.err
$TR bench.py --gpus $N --no-cpu --no-extras --steps 5 --warmup 3 > gpurun_out/r2t_bench_${N}gpu.json 2> gpurun_out/r2t_bench_${N}gpu.err
cat gpurun_out/r2t_pcie_${N}gpu.jsonl; tail -3 gpurun_out/r2t_pcie_${N}gpu.err; tail -3 gpurun_out/r2t_bench_${N}gpu.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2t_bench_${N}gpu.json"))
print("value", d["value"], "e2e", json.dumps(d["e2e"])[:1500])
print("f64", d["bit_exact_f64"]["value"], json.dumps(d["bit_exact_f64"]["e2e"])[:700])
h=json.load(open("gpurun_out/r2t_host_probe_${N}gpu.json")); print(h["cpu_count"], h["lscpu"], h["host_pack_unpack_144"].get("16")); print(h.get("topo"))
PY
