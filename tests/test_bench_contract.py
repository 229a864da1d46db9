"""bench.py's reference arm runs on the host alone (no GPU): the JSON line the driver parses must carry the contract's
keys for both kinds of baseline — the reference's own CPU code (cfg1, kind "reference" where oracle/_ref was built) and
the oracle port (kind "port") for the stages the reference does not have."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(args, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"] + args, capture_output=True, text=True,
                       timeout=600, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stderr
    return r.stdout


def test_reference_arm_legacy_workload(orc):
    out = _run(["--workload", "cfg1", "--steps", "1", "--warmup", "0"])
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1                                   # exactly one JSON line on stdout
    d = json.loads(lines[0])
    assert KEYS <= set(d) and d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == ("reference" if orc.have_ref() else "port") and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["higher_is_better"] is True


def test_reference_arm_chain_workload_and_nonzero_ranks_stay_silent():
    d = json.loads(_run(["--workload", "cfg2", "--steps", "1", "--warmup", "0"]).strip())
    assert KEYS <= set(d) and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert "256 samples x 128 chirps x 4 antennas" in d["config"]["workload"]
    # the GPU arm's workload keys under the GPU arm's names (the bounded CPU sample is named separately)
    assert d["config"]["frames_per_gpu_per_step"] == 1024 and d["config"]["sample_frames_per_step"] > 0
    # under torchrun only rank 0 works and prints
    assert _run(["--workload", "cfg2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2"}).strip() == ""
