"""Bare pinned host -> device copies on N GPUs AT THE SAME TIME (one process per GPU, torchrun), the size of one cfg3 bench
batch (402 653 184 bytes): what the host memory + PCIe fabric of the box carries, with none of our code on the path.
The e2e figure of bench.py at N GPUs cannot exceed aggregate / 6 291 456 bytes per frame.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 profiles/tools/h2d_probe_multi.py"""
import json
import os
import subprocess
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
n = 64 * 6291456
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h.fill_(1)
d = torch.empty(n, dtype=torch.uint8, device=dev)
pr = torch.cuda.get_device_properties(local)
bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
try:
    numa = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
except OSError:
    numa = None
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()


def timed(iters, together):
    if world > 1 and together:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return n * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9


together = timed(20, True)
# one rank at a time (the others idle): the per-link figure on the same box
alone = 0.0
for r in range(world):
    if world > 1:
        dist.barrier()
    if r == rank:
        alone = timed(10, False)
if world > 1:
    dist.barrier()
    t = torch.tensor([together, alone, float(numa if numa is not None else -1), float(pr.pci_bus_id)], device=dev, dtype=torch.float64)
    allv = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allv, t)
    rows = [v.cpu().tolist() for v in allv]
else:
    rows = [[together, alone, float(numa if numa is not None else -1), float(pr.pci_bus_id)]]
if rank == 0:
    agg = sum(r[0] for r in rows)
    out = {"probe": "pinned H2D, 402653184 bytes per copy, all ranks at once", "n_gpus": world,
           "per_gpu_gbs_together": [round(r[0], 1) for r in rows], "per_gpu_gbs_alone": [round(r[1], 1) for r in rows],
           "aggregate_gbs": round(agg, 1), "cfg3_frames_per_s_cap": round(agg * 1e9 / 6291456),
           "gpu_numa_node": [int(r[2]) for r in rows], "gpu_pci_bus": [f"{int(r[3]):02x}" for r in rows],
           "cpus_visible": len(os.sched_getaffinity(0))}
    print(json.dumps(out), flush=True)
    if os.environ.get("H2D_PROBE_TOPO"):
        try:
            print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout, flush=True)
        except Exception as e:  # noqa: BLE001
            print("nvidia-smi topo failed:", e)
if world > 1:
    dist.destroy_process_group()
