"""Static SASS opcode histogram of every kernel in libmmw_radar_b200.so (cuobjdump -sass): the Blackwell tell-tales first.
    python profiles/tools/sass_static.py > profiles/sass_r2.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "cuda-based-mmwave-radar-object-detection-acceleration_b200", "libmmw_radar_b200.so")
TELL = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "FADD2", "FFMA2", "FMUL2", "UTCMMA", "UTCHMMA", "LDTM", "STTM", "HMMA", "CCTL", "MEMBAR",
        "UCGABAR", "BAR", "LDS", "STS", "LDG", "STG", "ATOM", "RED", "ATOMS", "SHFL", "VOTE", "MUFU", "I2F", "I2FP", "F2I", "DFMA", "DADD", "DMUL"]

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip() or s
archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
kernels = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        kernels[cur][m.group(1)] += 1
total = collections.Counter()
for c in kernels.values():
    total.update(c)
print(f"# static SASS opcode counts of {os.path.basename(LIB)} (cuobjdump -sass); cubin architectures: {', '.join(archs)}")
print(f"# {len(kernels)} kernels, {sum(total.values())} instructions")
print("# whole library, tell-tale opcodes: " + ", ".join(f"{op} {total[op]}" for op in TELL if total[op]))
print("# absent: " + ", ".join(op for op in TELL if not total[op]))
print()
for name, c in kernels.items():
    n = sum(c.values())
    d = demangle(name)
    d = re.sub(r"\(mmw::PlanDev.*", "(...)", d)
    print(f"{d}\n    {n} instructions | " + ", ".join(f"{op} {c[op]}" for op in TELL if c[op]))
    print("    top: " + ", ".join(f"{op} {k}" for op, k in c.most_common(10)))
