"""Summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_list_summary.py <launches.csv> "<command that was profiled>"
Per kernel name: launches, total and average duration, share; then the same for the decode steps only (everything after the last
prompt compression)."""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"\(.*", "", name).replace("mfb::", "")
    return name.strip()


def table(rows, title):
    tot = sum(d for _, d in rows)
    agg = collections.OrderedDict()
    for n, d in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += d
    print(f"{title} ({len(rows)} launches):")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {c:5d} launches  {t:12.1f} us total  {100 * t / tot:5.1f} %  avg {t / c:9.1f} us  {n}")
    print()


def main():
    path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    lines = [l for l in open(path) if l.startswith('"')]
    rd = list(csv.reader(lines))
    hdr = rd[0]
    i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    rows = []
    for r in rd[1:]:
        v = float(r[i_val].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[i_unit], 1.0)
        rows.append((short(r[i_name]), v))
    print(cmd)
    print("(this library's kernels only; per-launch times under ncu are cold-cache and serialised: the SHARES are what compares with the bench)\n")
    table(rows, "all launches")
    last = max((i for i, (n, _) in enumerate(rows) if "compress_prefill" in n), default=-1)
    table(rows[last + 1:], "decode steps only (after the last compress_prefill: warm-up + timed steps of the device-resident and the end-to-end loop)")


if __name__ == "__main__":
    main()
