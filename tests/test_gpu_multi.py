"""Multi-GPU acceptance test of SURVEY.md §8e: the detection list gathered from N frame-sharded ranks over NCCL is
byte-identical to the single-GPU list.  Needs >= 2 GPUs (gpurun --gpus 2); skipped on a one-GPU box, where the same
logic is covered by the emulated-rank tests in test_gpu_properties.py and the gloo tests in test_sharding_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["nccl", "peer", "peer-fallback"])
def test_gather_equals_single_gpu(mode):
    """nccl: NCCL gather per step; peer: copy-engine puts into rank 0's memory (mmw_exchange_*), NCCL for the set-up only;
    peer-fallback: one rank cannot set the peer exchange up, every rank learns it at once and uses the NCCL gather"""
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", {"nccl": "29533", "peer": "29534", "peer-fallback": "29535"}[mode], os.path.join(ROOT, "tests", "dist_gather_check.py"), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "match=True" in r.stdout


def _cxx_example(tmp_path, name):
    pkg_dir = os.path.join(ROOT, "cuda-based-mmwave-radar-object-detection-acceleration_b200")
    exe = tmp_path / name
    r = subprocess.run(["/usr/bin/g++", "-O2", "-std=c++11", "-Wall", "-Werror", "-o", str(exe), os.path.join(ROOT, "examples", name + ".cpp"),
                        "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", "-L", pkg_dir, "-lmmw_radar_b200",
                        "-L", "/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + pkg_dir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_group_api_cxx_example(tmp_path):
    """examples/b200_group.cpp: a plain C++ host (include/ only) shards a batch over every visible GPU with mmw_group_*; the
    list gathered to GPU 0 over NCCL must equal the single-GPU list byte for byte.  On a one-GPU box the group has one
    member (no NCCL), which still exercises sharding rule, merged block and read-back."""
    exe = _cxx_example(tmp_path, "b200_group")
    r = subprocess.run([str(exe), "256", "128", "4", "11"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "identical=yes" in r.stdout
    print(r.stdout)


def test_group_api_python(pkg=None):
    """mmw_group_* through the ctypes mirror: host entry, device-resident shards (uneven, one empty), frame offsets"""
    import numpy as np
    import torch

    import __graft_entry__ as entry

    pkg = entry.load_package()
    n = torch.cuda.device_count()
    devices = list(range(min(n, 4)))
    S, C, A, F = 256, 128, 4, 9
    adc = pkg.synth.cube_batch(F, S, C, A, cfg=2, n_targets=6)
    with pkg.RadarContext(S, C, A, F, device=0) as whole:
        whole.set_frame_offset(100)
        want, _ = whole.process_host(adc, F)
    with pkg.api.RadarGroup(S, C, A, F, devices) as grp:           # capacity F per GPU: any split of the batch fits
        assert grp.size == len(devices)
        grp.set_frame_offset(100)
        got, ov = grp.process_host(adc, F)
        assert not ov and got.tobytes() == want.tobytes()
        # device-resident shards of uneven size, the last rank idle
        counts = [pkg.api.shard_frames(F, len(devices), r)[1] for r in range(len(devices))]
        if len(devices) > 1:
            counts[0] += counts[-1]
            counts[-1] = 0
        shards, first = [], 0
        for i, d in enumerate(devices):
            shards.append(torch.from_numpy(adc[first:first + counts[i]]).to(f"cuda:{d}") if counts[i] else None)
            first += counts[i]
        grp.process_device(shards, counts)
        again, ov = grp.read_detections()
        assert not ov and again.tobytes() == want.tobytes()
    for r in range(5):
        assert pkg.api.shard_frames(11, 5, r) == pkg.sharding.shard_frames(11, 5, r)
