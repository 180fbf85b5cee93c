"""Executed-instruction breakdown of an ncu report by SASS code region (runs of instructions with similar execution counts).
    python tools/ncu_regions.py <report.ncu-rep> [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ia, iex = hdr.index("Source"), hdr.index("Instructions Executed")
ist = hdr.index("Warp Stall Sampling (All Samples)")
data = []
for r in rows[2:]:
    if len(r) < len(hdr): break
    data.append((r[ia], int(r[iex]), int(r[ist] or 0)))
tot = sum(d[1] for d in data); tots = sum(d[2] for d in data)
seg, cur = [], None
for idx, (s, c, st) in enumerate(data):
    if cur is None or not (0.5 * cur['c'] <= c <= 2 * cur['c']):
        if cur: seg.append(cur)
        cur = {'start': idx, 'c': c, 'n': 0, 'sum': 0, 'st': 0, 'ops': {}}
    cur['n'] += 1; cur['sum'] += c; cur['st'] += st
    t = s.split(); op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    cur['ops'][op] = cur['ops'].get(op, 0) + 1
seg.append(cur)
seg.sort(key=lambda x: -x['sum'])
print(f"total executed {tot}, {len(data)} SASS instructions, stall samples {tots}")
for sg in seg[:top]:
    ops = sorted(sg['ops'].items(), key=lambda x: -x[1])[:8]
    print(f"idx {sg['start']:5d} n={sg['n']:4d} exec/instr~{sg['c']:9d} instr {sg['sum']/tot*100:5.1f}%  stall-samples {sg['st']/max(tots,1)*100:5.1f}%  {ops}")
