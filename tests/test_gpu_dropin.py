"""Drop-in proof: the reference's UNMODIFIED cudaBenchMarking.cpp (compiled from /root/reference into
oracle/_ref/cudaBenchMarking.o, linked against libmmw_radar_b200.so) runs its cpuTiming()/cudaTiming()
loops on a synthetic fhy_direct.bin."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu


def test_unmodified_reference_caller_runs_against_our_library(pkg, tmp_path):
    exe = pkg.build.DROPIN_BIN
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin_acceleration was not built (needs /root/reference at build time)")
    cap = pkg.synth.legacy_capture(90, seed=0)
    cap.tofile(tmp_path / "fhy_direct.bin")
    assert os.path.getsize(tmp_path / "fhy_direct.bin") == 18_432_000
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    assert out.count("Inner CUDA Timing") == 89                     # one line per frame, as acceleration.cu:533
    assert "Total Time for 89 frames" in out and "cuda totalTime" in out and "cuda inner time" in out
    print("\n" + "\n".join(l for l in out.splitlines() if "Inner CUDA" not in l))
