"""Summarise an ncu --set full report of sparse_decode_attn_kernel into markdown (run where ncu is installed).
    python tools/ncu_summary.py <report.ncu-rep> <tiles> <title>
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep, tiles, title = sys.argv[1], float(sys.argv[2]), sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, d = rows[0], rows[1], dict(zip(rows[0], rows[2]))
    print(f"## {title}\n\nreport: `{rep}` (ncu --set full --clock-control none --import-source on), kernel `{d.get('Kernel Name')}`\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in d:
            print(f"| {k} | {d[k]} | {units[hdr.index(k)]} |")
    print(f"| instructions per 64-position tile | {float(d['smsp__inst_executed.sum']) / tiles:.2f} | |")
    print(f"| shared-memory wavefronts per tile | {float(d['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']) / tiles:.2f} | |")
    print(f"| SM-active clocks per tile per SM | {float(d['sm__cycles_active.avg']) * 148 / tiles:.2f} | |")
    print("\nwarp stall reasons (warps per issue-active cycle): " + ", ".join(
        f"{k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {float(d[k]):.2f}"
        for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(d[k]) > 0.1))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    ia, iex = hdr.index("Source"), hdr.index("Instructions Executed")
    op = collections.Counter()
    for r in rows[2:]:
        if len(r) < len(hdr):
            break
        toks = r[ia].split()
        name = toks[1] if toks[0].startswith("@") else toks[0]
        op[name.split(".")[0]] += int(r[iex])
    print("\nSASS opcodes executed per tile: " + ", ".join(f"{k} {v / tiles:.2f}" for k, v in op.most_common(16)))
    print()


if __name__ == "__main__":
    main()
