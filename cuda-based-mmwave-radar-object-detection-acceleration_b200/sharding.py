"""Frame sharding across the GPUs of one box and the gather of detection lists to rank 0.

Frames are independent (in the reference the only state shared between frames is
the read-only base frame, cudaBenchMarking.cpp:374-378), so a batch is split into contiguous
frame blocks, one per rank, and every rank runs the whole chain on its own block with no
data-path collective.  The single exchange step is the gather of the variable-length detection
lists (24-byte records, KBs per batch) to rank 0: one all_gather of the per-rank counts and one
point-to-point send per non-empty rank.  Because rank r owns frames [first_r, first_r + n_r) and
each local list is already ordered by (frame, range, doppler), concatenating in rank order gives
the globally ordered list with no sort.

Works with the nccl backend (CUDA tensors) and with gloo (CPU tensors; used by the CPU tests).
"""
from __future__ import annotations

import os

import numpy as np

REC_BYTES = 24


def shard_frames(n_frames: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous block of frames owned by `rank`: (first_frame, count). Remainders go to the low ranks."""
    base, rem = divmod(n_frames, world_size)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


class _CudaView:
    """Exposes a raw device allocation owned by libmmw_radar_b200.so through the CUDA array
    interface so that torch can wrap it without a copy (torch.as_tensor(view, device='cuda'))."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {
            "shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None,
        }


def device_bytes_view(ptr: int, nbytes: int, device):
    import torch

    return torch.as_tensor(_CudaView(ptr, nbytes), device=device)


def gather_detections(local_records, n_local, group=None):
    """Gathers per-rank detection lists to rank 0.

    local_records : torch.uint8 tensor [cap * 24] (CUDA for nccl, CPU for gloo) holding n_local records
    n_local       : torch.int64 tensor [1] on the same device (number of valid records)
    returns       : on rank 0 a torch.uint8 tensor [total * 24] in rank order (frames ascending); None elsewhere
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = torch.empty(world, dtype=torch.int64, device=n_local.device)
    dist.all_gather_into_tensor(counts, n_local, group=group)
    counts_h = counts.cpu().tolist()                       # the one host sync of the exchange step
    mine = counts_h[rank]
    if rank != 0:
        if mine > 0:
            dist.send(local_records[: mine * REC_BYTES], dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return None
    total = sum(counts_h)
    out = torch.empty(total * REC_BYTES, dtype=torch.uint8, device=local_records.device)
    out[: mine * REC_BYTES].copy_(local_records[: mine * REC_BYTES])
    off = mine
    for r in range(1, world):
        if counts_h[r] > 0:
            src = dist.get_global_rank(group, r) if group is not None else r
            dist.recv(out[off * REC_BYTES: (off + counts_h[r]) * REC_BYTES], src=src, group=group)
            off += counts_h[r]
    return out


class DetectionGather:
    """The exchange step of the sharded chain: no host round trip, and off the compute stream.

    Every rank contributes a FIXED-size prefix of its contiguous result block ([32-byte header | ordered
    records], mmw_device_result_block) — `records_per_rank` records, a few hundred KB — so the sizes NCCL needs
    are known up front and nothing has to be copied to the host first.  Rank 0 receives the blocks in rank order
    with one gather over NVLink and packs them into one ordered list with one launch of the library's merge
    kernel (mmw_merge_gathered), reading the true counts from the gathered headers on the device.  A rank that
    produced more than `records_per_rank` detections is truncated and the merged header's overflow word is set.

    Overlap: run() snapshots the result block on the compute stream (one small device copy; the next batch
    overwrites the block) and hands everything else — the NCCL gather and the merge kernel — to a side stream
    that only depends on that snapshot, so the exchange of step k runs under the kernels of step k+1.  Send,
    receive and merged buffers are double-buffered; flush() makes the compute stream wait for the last exchange.
    The context's stream must be torch's current stream when run() is called."""

    def __init__(self, ctx, device, records_per_rank: int, group=None, side=None):
        import torch
        import torch.distributed as dist

        self.torch = torch
        self.ctx, self.group = ctx, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        block, cap_bytes = ctx.device_result_block()
        self.stride = 32 + REC_BYTES * records_per_rank
        if self.stride > cap_bytes:
            raise ValueError("records_per_rank exceeds the context's detection capacity")
        self.local = device_bytes_view(block, self.stride, device)
        self.send = [torch.empty(self.stride, dtype=torch.uint8, device=device) for _ in range(2)]
        self.merged_cap = self.world * records_per_rank
        # several gathers (one per batch in flight) must share ONE side stream: NCCL calls on a communicator stay in issue order
        self.side = side if side is not None else torch.cuda.Stream(device=device)
        self.snap = [torch.cuda.Event() for _ in range(2)]       # snapshot k is in send[k]
        self.done = [None, None]                                   # exchange that last used buffer pair k has finished
        self.step = 0
        self.latest = None
        if self.rank == 0:
            self.gathered = [torch.empty((self.world, self.stride), dtype=torch.uint8, device=device) for _ in range(2)]
            self.slots = [list(g.unbind(0)) for g in self.gathered]
            self.merged = [torch.empty(32 + REC_BYTES * self.merged_cap, dtype=torch.uint8, device=device) for _ in range(2)]

    def run(self):
        """call after ctx.process_device() with the context's stream current.  Returns the merged block of this
        step on rank 0 (None elsewhere); it is complete once flush() has been called or the side stream has been
        waited for."""
        import torch.distributed as dist

        torch = self.torch
        k = self.step & 1
        self.step += 1
        compute = torch.cuda.current_stream()
        dbg = os.environ.get("MMW_GATHER_DEBUG", "full")          # experiment switch: which parts of the exchange run
        if dbg != "full" and self.rank == 0:
            self.latest = self.merged[k]                          # (not a result: the timing experiment only)
        if dbg == "none":
            return self.latest if self.rank == 0 else None
        if self.done[k] is not None:
            compute.wait_event(self.done[k])                      # two steps old: never actually waits
        self.send[k].copy_(self.local)
        self.snap[k].record(compute)
        dst = dist.get_global_rank(self.group, 0) if self.group is not None else 0
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.snap[k])
            if dbg != "snap":
                dist.gather(self.send[k], gather_list=self.slots[k] if self.rank == 0 else None, dst=dst, group=self.group)
            if self.rank == 0 and dbg == "full":
                self.ctx.use_stream(self.side.cuda_stream)
                self.ctx.merge_gathered(self.gathered[k], self.world, self.stride, self.merged[k], self.merged_cap)
                self.ctx.use_stream(compute.cuda_stream)
                self.latest = self.merged[k]
            ev = torch.cuda.Event()
            ev.record(self.side)
            self.done[k] = ev
        return self.latest if self.rank == 0 else None

    def flush(self):
        """the current stream waits for every exchange started so far; returns the last merged block on rank 0"""
        cur = self.torch.cuda.current_stream()
        for ev in self.done:
            if ev is not None:
                cur.wait_event(ev)
        return self.latest if self.rank == 0 else None

    def read(self, det_dtype):
        """rank 0: (records, header) of the last merged block on the host (synchronises)"""
        self.flush()
        m = self.latest.cpu().numpy()
        header = m[:32].view(np.uint32).copy()
        n = int(header[0])
        return np.frombuffer(m[32:32 + REC_BYTES * n].tobytes(), dtype=det_dtype), header


class PeerExchangeUnavailable(RuntimeError):
    """mmw_exchange_* could not be set up on at least one rank (raised on every rank of the group at once)"""


class PeerDetectionGather:
    """The exchange step without a kernel on the data path (mmw_exchange_*, include/mmw_radar.h): every rank's result block
    goes straight into rank 0's memory over NVLink by a copy-engine put, flags and credits are stream memory operations, and
    rank 0 runs the one merge kernel on a side stream.  torch.distributed (NCCL) only carries the 64-byte IPC handles at
    set-up.  Same interface as DetectionGather: run() after ctx.process_device() on the context's stream, flush(), read()."""

    def __init__(self, ctx, device, records_per_rank: int, group=None, depth: int = 4):
        import torch
        import torch.distributed as dist

        from .api import PeerExchange

        self.torch, self.ctx, self.device = torch, ctx, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.records_per_rank = records_per_rank

        def all_gather_bytes(b: bytes):
            mine = torch.tensor(list(b), dtype=torch.uint8, device=device)
            everyone = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(everyone, mine, group=group)
            return [bytes(t.cpu().tolist()) for t in everyone]

        self.x = PeerExchange(ctx, self.rank, self.world, records_per_rank, all_gather_bytes, depth=depth)
        # every rank is connected before the first put - or none uses the exchange: the reduction is the barrier
        ok = torch.tensor([0 if self.x.error is not None else 1], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            why = str(self.x.error) if self.x.error is not None else "another rank could not set the exchange up"
            self.x.close()
            raise PeerExchangeUnavailable(why)                     # raised on EVERY rank: the caller falls back collectively
        self.merged_bytes = 32 + REC_BYTES * self.world * records_per_rank
        self.latest = None

    def run(self):
        self.x.put()
        if self.rank == 0:
            self.latest = self.x.merge()
        return self.latest

    def flush(self):
        """the context's stream waits for the last merge (rank 0); returns the merged block's device pointer"""
        if self.rank == 0 and self.latest is not None:
            self.latest = self.x.wait(self.ctx.stream)
        return self.latest

    def read(self, det_dtype):
        """rank 0: (records, header) of the last merged block on the host (synchronises)"""
        self.flush()
        self.torch.cuda.synchronize(self.device)
        m = device_bytes_view(self.latest, self.merged_bytes, self.device).cpu().numpy()
        header = m[:32].view(np.uint32).copy()
        n = int(header[0])
        return np.frombuffer(m[32:32 + REC_BYTES * n].tobytes(), dtype=det_dtype), header

    def close(self):
        self.x.close()


def records_from_bytes(buf, det_dtype) -> np.ndarray:
    """torch.uint8 tensor (any device) -> numpy structured array of detections."""
    return np.frombuffer(buf.cpu().numpy().tobytes(), dtype=det_dtype)
