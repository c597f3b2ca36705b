"""Per-source-line totals of an ncu source page: python tools/ncu_lines.py file.ncu-rep object.o mangled-kernel-substring [min_pct]
The SASS rows of `ncu --page source --csv` are matched, in order, with `nvdisasm -g` of the same object (line markers)."""
import csv, os, re, subprocess, sys, tempfile
rep, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
minpct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":")][0]
lines = []          # (line number of the innermost frame, outermost line) per instruction
cur = None
for l in dis[start + 1:]:
    if l.startswith("//-----") or (l.startswith(".text.") and l.rstrip().endswith(":")):
        break
    mm = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if mm:
        cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
assert len(data) == len(lines), (len(data), len(lines))
iins, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
for r, ln in zip(data, lines):
    a = agg.setdefault(ln, [0.0, 0.0, {}])
    a[0] += float(r[iins] or 0); a[1] += float(r[ismp] or 0)
    for i in stall:
        v = float(r[i] or 0)
        if v: a[2][hdr[i]] = a[2].get(hdr[i], 0) + v
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("total warp-instructions %.4g, samples %d" % (ti, ts))
srcs = {}
for (f, n), a in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if a[0] / ti * 100 < minpct and a[1] / ts * 100 < minpct:
        continue
    if f not in srcs:
        for d in ("qldpc_b200/csrc", "."):
            pth = os.path.join(d, f)
            if os.path.exists(pth):
                srcs[f] = open(pth).read().splitlines(); break
        else:
            srcs[f] = []
    text = srcs[f][n - 1].strip()[:100] if n - 1 < len(srcs[f]) else ""
    top = sorted(a[2].items(), key=lambda kv: -kv[1])[:2]
    print("%s:%-5d %5.1f%% inst %5.1f%% smp  [%s]  %s" % (f, n, a[0] / ti * 100, a[1] / ts * 100, ", ".join("%s %.0f%%" % (k[6:], v / max(a[1], 1) * 100) for k, v in top), text))
