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


def records_from_bytes(buf, det_dtype) -> np.ndarray:
    """torch.uint8 tensor (any device) -> numpy structured array of detections."""
    return np.frombuffer(buf.cpu().numpy().tobytes(), dtype=det_dtype)
