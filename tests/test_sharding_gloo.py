"""Multi-rank host logic on CPU: frame sharding and the detection-list gather, world_size 2 and 3 over gloo."""
import os
import socket

import numpy as np
import pytest

import __graft_entry__ as entry


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_frames_partitions_exactly():
    pkg = entry.load_package()
    for n in (0, 1, 7, 64, 1024, 1025):
        for w in (1, 2, 3, 4, 8):
            spans = [pkg.sharding.shard_frames(n, w, r) for r in range(w)]
            assert sum(c for _, c in spans) == n
            nxt = 0
            for first, cnt in spans:
                assert first == nxt and cnt in (n // w, n // w + 1)
                nxt += cnt


def _worker(rank, world, port, n_frames, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = entry.load_package()
    orc = entry.load_oracle()
    S, C, A = 64, 64, 2
    first, cnt = pkg.sharding.shard_frames(n_frames, world, rank)
    wr, wd = orc.hann_periodic(S), orc.hann_periodic(C)
    # each rank produces the records of ITS frames (the CPU oracle stands in for the GPU chain here;
    # what is under test is the sharding + gather plumbing)
    if cnt:
        adc = pkg.synth.cube_batch(cnt, S, C, A, cfg=4, first_frame=first, n_targets=2)
        dets = orc.process_frames(adc, cnt, S, C, A, wr, wd)["dets"]
        dets["frame"] += first
    else:
        dets = np.zeros(0, orc.DET_DTYPE)
    buf = torch.zeros(max(1, len(dets)) * 24 + 240, dtype=torch.uint8)      # capacity > valid part
    buf[: len(dets) * 24] = torch.frombuffer(bytearray(dets.tobytes()), dtype=torch.uint8) if len(dets) else buf[:0]
    out = pkg.sharding.gather_detections(buf, torch.tensor([len(dets)], dtype=torch.int64))
    if rank == 0:
        q.put(pkg.sharding.records_from_bytes(out, orc.DET_DTYPE))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames", [(2, 5), (3, 2)])
def test_gather_equals_single_rank(world, n_frames):
    import torch.multiprocessing as mp

    pkg = entry.load_package()
    orc = entry.load_oracle()
    orc.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    S, C, A = 64, 64, 2
    adc = pkg.synth.cube_batch(n_frames, S, C, A, cfg=4, n_targets=2)
    want = orc.process_frames(adc, n_frames, S, C, A, orc.hann_periodic(S), orc.hann_periodic(C))["dets"]
    assert len(want) > 0 and got.tobytes() == want.tobytes()      # byte-identical to the 1-rank list


def test_c_abi_shard_rule_is_the_python_rule():
    """mmw_shard_frames (the rule mmw_group_* uses inside the library) and sharding.shard_frames (the torchrun launchers' rule)
    are the same function: contiguous blocks, remainders to the low ranks — pure host code, no GPU"""
    import __graft_entry__ as entry

    pkg = entry.load_package()
    for n in (0, 1, 7, 11, 64, 1000, 4097):
        for w in (1, 2, 3, 5, 8):
            spans = [pkg.api.shard_frames(n, w, r) for r in range(w)]
            assert spans == [pkg.sharding.shard_frames(n, w, r) for r in range(w)]
            assert sum(c for _, c in spans) == n and all(spans[r][0] + spans[r][1] == spans[r + 1][0] for r in range(w - 1))
