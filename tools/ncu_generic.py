"""Generic markdown summary of an ncu --set full report (first kernel in the report): python tools/ncu_generic.py <rep> <title> [units per launch] [unit name]"""
import collections, csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
rep, title = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else 0
uname = sys.argv[4] if len(sys.argv) > 4 else "unit"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, un, d = rows[0], rows[1], dict(zip(rows[0], rows[2]))
print(f"## {title}\n\nreport: `{rep}` (ncu --set full --clock-control none --import-source on), kernel `{d.get('Kernel Name')}`\n")
print("| metric | value | unit |\n|---|---|---|")
for k in KEYS:
    if k in d:
        print(f"| {k} | {d[k]} | {un[hdr.index(k)]} |")
if units:
    print(f"| instructions per {uname} | {float(d['smsp__inst_executed.sum']) / units:.1f} | |")
    print(f"| SM-active clocks per {uname} per SM | {float(d['sm__cycles_active.avg']) * 148 / units:.1f} | |")
print("\nwarp stall reasons (warps per issue-active cycle): " + ", ".join(
    f"{k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {float(d[k]):.2f}"
    for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(d[k]) > 0.1))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h2 = rows[1]
ia, iex = h2.index("Source"), h2.index("Instructions Executed")
op = collections.Counter()
for r in rows[2:]:
    if len(r) < len(h2):
        break
    t = r[ia].split()
    op[(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += int(r[iex])
tot = sum(op.values())
print("\nSASS opcodes (share of executed warp instructions): " + ", ".join(f"{k} {v / tot * 100:.1f} %" for k, v in op.most_common(14)))
