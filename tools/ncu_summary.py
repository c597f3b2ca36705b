"""Summarise an .ncu-rep (raw page + source page) into text: python tools/ncu_summary.py file.ncu-rep [min_pct] [kernel-regex | name:invocation]
(name:invocation selects the n-th launch of a kernel name, e.g. bp_stage_kernel:3 -- template arguments are not part of the name)"""
import csv, subprocess, sys
rep = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
ksel = []
if len(sys.argv) > 3:
    ksel = ["--kernel-id", "::" + sys.argv[3]] if ":" in sys.argv[3] else ["--kernel-name", "regex:" + sys.argv[3]]
raw = subprocess.run(["ncu", "-i", rep] + ksel + ["--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active"]
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k:85s} {vals[i][:60]:>22s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep] + ksel + ["--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
data = []
for r in rows[2:]:
    if r == hdr or (r and r[0] == 'Kernel Name'):
        break              # a second view / kernel follows
    if len(r) == len(hdr):
        data.append(r)
isrc, ins, ismp, ithr = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
iwf = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
iwfi = hdr.index("L1 Wavefronts Shared Ideal") if "L1 Wavefronts Shared Ideal" in hdr else None
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ti = sum(int(r[ins]) for r in data); ts = sum(int(r[ismp]) for r in data)
print(f"\ntotal warp-instructions {ti}  samples {ts}")
tot = {hdr[i]: sum(int(r[i]) for r in data) for i in stall}
print("stall mix: " + ", ".join(f"{k[6:]} {100*v/ts:.1f}%" for k, v in sorted(tot.items(), key=lambda x: -x[1])[:8]))
if iwf is not None:
    wf = sum(int(r[iwf]) for r in data); wfi = sum(int(r[iwfi]) for r in data)
    print(f"shared wavefronts {wf} ideal {wfi} (x{wf/max(1,wfi):.2f})")
print(f"\ninstructions with >= {minpct}% of executed warp-instructions:")
for idx, r in enumerate(data):
    n = int(r[ins])
    if n >= minpct / 100 * ti:
        top = sorted(((int(r[i]), hdr[i][6:]) for i in stall), reverse=True)[:2]
        w = f"wf {r[iwf]}/{r[iwfi]}" if (iwf is not None and int(r[iwf])) else ""
        print(f"{idx:5d} {r[isrc].strip()[:58]:58s} inst {100*n/ti:5.2f}% smp {100*int(r[ismp])/ts:5.2f}% thr {r[ithr]:>4s} {top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]} {w}")
