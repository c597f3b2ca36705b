"""Host-side probe of the GPU box: CPU model / core count / NUMA nodes, and the throughput of the host pack / unpack
routines (qldpc_b200/csrc/host_pack.h, built with g++ on the spot) for 1 .. N threads."""
import ctypes, os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
out["cpu_count"] = os.cpu_count()
out["affinity"] = len(os.sched_getaffinity(0))
try:
    out["lscpu"] = {l.split(":")[0].strip(): l.split(":", 1)[1].strip() for l in subprocess.run(["lscpu"], capture_output=True, text=True).stdout.splitlines()
                    if l.split(":")[0].strip() in ("Model name", "Socket(s)", "Core(s) per socket", "Thread(s) per core", "NUMA node(s)", "CPU(s)", "Flags", "NUMA node0 CPU(s)", "NUMA node1 CPU(s)")}
    fl = out["lscpu"].pop("Flags", "")
    out["flags"] = [f for f in ("sse2", "avx2", "bmi2", "avx512f", "avx512bw", "avx512vbmi") if f in fl.split()]
except Exception as e:
    out["lscpu"] = str(e)
try:
    out["meminfo"] = open("/proc/meminfo").read().splitlines()[:3]
    out["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
except Exception as e:
    pass
so = "/tmp/hp_probe.so"
subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", so, os.path.join(ROOT, "tests", "cpp", "host_pack_check.cpp")],
               check=True, env={k: v for k, v in os.environ.items() if k not in ("CC", "CXX")})
L = ctypes.CDLL(so)
L.hp_time.restype = ctypes.c_double
rates = {}
for t in (1, 2, 4, 8, 12, 16, 24, 32, 48, 64):
    if t > out["affinity"]:
        break
    dt = L.hp_time(ctypes.c_longlong(1 << 20), 72, 144, t, 5)
    rates[t] = dict(ms_per_2_20_shots=round(dt * 1e3, 3), shots_per_s=(1 << 20) / dt, gbytes_per_s=(1 << 20) * (72 + 144 + 12 + 20) / dt / 1e9)
out["host_pack_unpack_144"] = rates
print(json.dumps(out, indent=1))
