"""Summarise an .ncu-rep here (no GPU): key raw metrics, top source lines, opcode mix."""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
which = int(sys.argv[3]) if len(sys.argv) > 3 else None   # index of the kernel's source section (default: the longest)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, data = rows[0], rows[1], rows[2:]
keys = ['Kernel Name','gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','sm__warps_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__warps_eligible.avg.per_cycle_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
for k in keys:
    if k in h:
        i = h.index(k); print(f"{k:72s} {units[i]:10s} {[r[i][:60] for r in data]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdrs = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
best = None
for hi, hd in enumerate(hdrs):
    end = hdrs[hi + 1] - 3 if hi + 1 < len(hdrs) else len(rows)
    sec = rows[hd + 1:end]
    if (which is None and (best is None or len(sec) > len(best[1]))) or which == hi: best = (rows[hd], sec, rows[hd - 2][:2] if hd >= 2 else None)
hh, sec, fn = best
print("source file:", fn)
iS, iI = hh.index("# Samples"), hh.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(hh) if c.startswith("stall_") and "Not Issued" not in c]
agg, ops, stalls = [], Counter(), Counter()
cur = None
for r in sec:
    if len(r) <= iI: continue
    if r[0].strip().isdigit():
        cur = (int(r[0]), r[1].strip()[:95])
        try: agg.append((int(r[iI]), int(r[iS]), cur[0], cur[1]))
        except ValueError: pass
        for i in stall_cols:
            try: stalls[hh[i]] += int(r[i])
            except ValueError: pass
    elif len(r) > 3 and r[2].startswith("0x"):
        try: ops[(r[3].strip().split() or [''])[0].split('.')[0]] += int(r[iI])
        except ValueError: pass
tot, tots = sum(a[0] for a in agg), sum(a[1] for a in agg)
print("total warp-inst", tot, "samples", tots)
for a in sorted(agg, reverse=True)[:topn]:
    print(f"{100*a[0]/max(tot,1):5.1f}% inst {100*a[1]/max(tots,1):5.1f}% smp  L{a[2]:4d}: {a[3]}")
st = sum(ops.values())
print("opcode mix:", [(k, round(100 * v / st, 1)) for k, v in ops.most_common(18)])
ss = sum(stalls.values())
print("stall mix:", [(k, round(100 * v / ss, 1)) for k, v in stalls.most_common(8)])
