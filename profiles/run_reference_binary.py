"""the UNMODIFIED reference program (oracle/_ref/ref_acceleration: its cudaBenchMarking.cpp +
acceleration.cu compiled for sm_100a by oracle/Makefile) and the same unmodified caller linked against our library
(oracle/_ref/dropin_acceleration), both on the same synthetic fhy_direct.bin (90 frames, reference format), on this box.
Prints the summary lines each program prints (the per-frame "Inner CUDA Timing" lines are dropped)."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
with tempfile.TemporaryDirectory() as d:
    pkg.synth.legacy_capture(90, seed=0).tofile(os.path.join(d, "fhy_direct.bin"))
    for name in ("ref_acceleration", "dropin_acceleration"):
        exe = os.path.join(ROOT, "oracle", "_ref", name)
        if not os.path.exists(exe):
            print(f"{name}: not built (needs /root/reference at build time)")
            continue
        for rep in range(2):                      # second run: warm page cache / driver
            r = subprocess.run([exe], cwd=d, capture_output=True, text=True, timeout=600)
        print(f"===== {name} (exit {r.returncode}) =====")
        print("\n".join(l for l in r.stdout.splitlines() if "Inner CUDA" not in l))
        if r.stderr.strip():
            print("stderr:", r.stderr.strip()[:500])

probe = os.path.join(ROOT, "profiles", "tools", "init_probe")          # g++ -O2 init_probe.cpp (see its header), run from the repo root
if os.path.exists(probe):
    print("\n===== profiles/tools/init_probe (same box): where the first call's time goes =====")
    for rep in range(2):
        r = subprocess.run([probe], cwd=ROOT, capture_output=True, text=True, timeout=600)
        print(r.stdout.strip() + ("   (second process on the box)" if rep else ""))
