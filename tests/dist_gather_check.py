"""Worker of tests/test_gpu_multi.py: run under torchrun with one rank per GPU.  Every rank processes its contiguous
frame block of the same seeded batch, the detection lists are gathered to rank 0 — argv[1] = "nccl": sharding.DetectionGather
(NCCL gather + merge kernel), "peer": sharding.PeerDetectionGather (copy-engine puts into rank 0's memory, mmw_exchange_*),
"peer-fallback": the peer exchange refused by one rank, all ranks falling back to the NCCL gather together —
over several pipelined steps (more than the exchange's ring depth), and rank 0 compares the merged list with the list one
GPU computes for the whole batch.  Exit code 0 = byte-identical."""
import faulthandler
import os
import sys

faulthandler.enable()

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    pkg = entry.load_package()
    S, C, A, F = 256, 128, 4, 11                                  # 11 frames: uneven shards
    adc = pkg.synth.cube_batch(F, S, C, A, cfg=2, n_targets=6)
    first, cnt = pkg.sharding.shard_frames(F, world, rank)
    ctx = pkg.RadarContext(S, C, A, max(cnt, 1), max_det_per_frame=2048, device=local)       # capacity >= the exchange's 2048 records on every rank
    ctx.set_frame_offset(first)
    stream = torch.cuda.Stream(device=dev)
    ctx.use_stream(stream.cuda_stream)
    mine = torch.from_numpy(adc[first:first + cnt]).to(dev)
    mode = sys.argv[1] if len(sys.argv) > 1 else "nccl"
    if mode == "peer-fallback":                                   # one rank cannot set the peer exchange up: every rank must see
        os.environ["MMW_EXCHANGE_FAIL_RANK"] = str(world - 1)     # PeerExchangeUnavailable and switch to the NCCL gather together
        try:
            pkg.sharding.PeerDetectionGather(ctx, dev, 2048)
            raise SystemExit("the peer exchange came up although a rank refused it")
        except pkg.sharding.PeerExchangeUnavailable:
            gather = pkg.sharding.DetectionGather(ctx, dev, 2048)
    else:
        gather = pkg.sharding.PeerDetectionGather(ctx, dev, 2048) if mode == "peer" else pkg.sharding.DetectionGather(ctx, dev, 2048)
    with torch.cuda.stream(stream):
        for _ in range(11):                                       # several steps: exercises the double buffering / the slot ring
            ctx.process_device(mine, cnt)
            gather.run()
        gather.flush()
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        recs, hdr = gather.read(pkg.DET_DTYPE)
        with pkg.RadarContext(S, C, A, F, max_det_per_frame=2048, device=local) as whole:
            want, _ = whole.process_host(adc, F)
        ok = recs.tobytes() == want.tobytes() and int(hdr[0]) == len(want) and int(hdr[2]) == F and int(hdr[3]) == 0
        print(f"world {world} ({mode}): {len(recs)} gathered detections, match={ok}", flush=True)
    dist.barrier()
    if mode == "peer":
        gather.close()
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
