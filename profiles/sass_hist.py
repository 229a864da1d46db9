"""Opcode histogram + top stall lines from an `ncu --page source --csv` dump (one kernel).
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME > k.csv ; python profiles/sass_hist.py k.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter()
stall = []
tot = 0
samples = 0
for r in rows[2:]:
    if r and r[0] == "Address":
        continue
    if len(r) < len(hdr):
        continue
    sass = r[idx["Source"]].strip()
    n = int(r[idx["Instructions Executed"]] or 0)
    s = int(r[idx["# Samples"]] or 0)
    toks = sass.split()
    op = toks[0] if not toks[0].startswith("@") else toks[1]
    op = op.split(".")[0] + ("." + ".".join(op.split(".")[1:2]) if op.startswith(("LDS", "STS", "LDG", "STG")) else "")
    ops[op] += n
    tot += n
    samples += s
    stall.append((s, n, sass))
print(f"total warp instructions {tot}, stall samples {samples}")
for op, n in ops.most_common(28):
    print(f"  {op:14s} {n:12d} {100.0 * n / tot:6.2f}%")
print("top stall lines:")
for s, n, sass in sorted(stall, reverse=True)[:25]:
    print(f"  {s:6d} samples  {n:9d} exec  {sass}")
