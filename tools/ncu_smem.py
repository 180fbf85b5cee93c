"""Shared-memory wavefronts per SASS instruction of an ncu report (top offenders + totals by opcode).
    python tools/ncu_smem.py <report.ncu-rep> [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
c = {n: hdr.index(n) for n in ("Source", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "L1 Wavefronts Shared Excessive")}
data = []
for i, r in enumerate(rows[2:]):
    if len(r) < len(hdr): break
    wf = int(r[c["L1 Wavefronts Shared"]] or 0)
    if wf: data.append((i, r[c["Source"]], int(r[c["Instructions Executed"]]), wf, int(r[c["L1 Wavefronts Shared Ideal"]] or 0)))
tot = sum(d[3] for d in data); ideal = sum(d[4] for d in data)
print(f"shared wavefronts {tot}, ideal {ideal}, excess {tot - ideal} ({(tot - ideal) / max(tot, 1) * 100:.1f} %)")
byop = {}
for i, s, ex, wf, idl in data:
    t = s.split(); op = ".".join((t[1] if t[0].startswith('@') else t[0]).split('.')[:3])
    a = byop.setdefault(op, [0, 0, 0]); a[0] += ex; a[1] += wf; a[2] += idl
for op, (ex, wf, idl) in sorted(byop.items(), key=lambda x: -x[1][1]):
    print(f"  {op:24s} exec {ex:10d} wavefronts {wf:10d} ({wf / tot * 100:5.1f} %) ideal {idl:10d}  wf/instr {wf / max(ex, 1):.2f}")
print("top instructions by excess wavefronts:")
for i, s, ex, wf, idl in sorted(data, key=lambda d: -(d[3] - d[4]))[:top]:
    print(f"  idx {i:5d} exec {ex:9d} wf {wf:9d} ideal {idl:9d}  {s[:90]}")
