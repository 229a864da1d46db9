"""Summarises an `ncu --metrics gpu__time_duration.sum,dram__bytes_* --csv` launch list: per-kernel launches,
mean duration, DRAM bytes per launch and share of the summed device time."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows[1:]:
    key = (r[ix["ID"]], r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("mmw::", ""))
    per.setdefault(key, {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
agg = collections.OrderedDict()
for (_, name), d in per.items():
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
total = sum(a[1] for a in agg.values())
print(f"| kernel | launches | mean us | share of device time | DRAM read MB/launch | DRAM write MB/launch |")
print("|---|---|---|---|---|---|")
for name, (n, t, rd, wr) in agg.items():
    print(f"| `{name}` | {n} | {t / n / 1e3:.1f} | {100 * t / total:.1f} % | {rd / n / 1e6:.1f} | {wr / n / 1e6:.1f} |")
