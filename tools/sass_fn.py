"""Extract one kernel's SASS from `cuobjdump -sass` output and histogram its opcodes (whole function and the hottest
backward-branch loop).  usage: python tools/sass_fn.py all.sass <substring of the mangled name> [--dump]"""
import re, sys, collections
src, key = sys.argv[1], sys.argv[2]
lines = open(src).read().split("\n")
start = [i for i, l in enumerate(lines) if "Function :" in l and key in l]
if not start:
    sys.exit("not found")
s = start[0]
e = next((i for i in range(s + 1, len(lines)) if "Function :" in lines[i]), len(lines))
ins = []
for l in lines[s:e]:
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(3), l.strip()))
print(lines[s].strip(), len(ins), "instructions")
def hist(sub):
    c = collections.Counter(op.split(".")[0] for _, op, _ in sub)
    return " ".join("%s:%d" % kv for kv in c.most_common())
print("ALL:", hist(ins))
# loops = backward branches
loops = []
for a, op, l in ins:
    if op.startswith("BRA"):
        m = re.search(r"0x([0-9a-f]+)", l.split("BRA")[1])
        if m and int(m.group(1), 16) < a:
            loops.append((int(m.group(1), 16), a))
for lo, hi in sorted(loops, key=lambda t: t[0] - t[1])[:4]:
    sub = [x for x in ins if lo <= x[0] <= hi]
    print("LOOP %x-%x (%d):" % (lo, hi, len(sub)), hist(sub))
if "--dump" in sys.argv:
    for _, _, l in ins: print(l)
