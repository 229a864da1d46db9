#!/usr/bin/env python
"""bench.py — radar frames/s (ADC cube -> detections) on N B200s, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2] [--impl ours|reference]

A step = one pass of the whole chain (range FFT -> Doppler FFT + |X|^2 integration -> 2-D CA-CFAR ->
detection records incl. angle FFT + grouping -> dense list) over one batch of synthetic frames per GPU.
  value  frames/s with the batch already resident in HBM (device-timed, max over ranks; for N > 1 the
         timed step includes the NCCL gather of the detection lists to rank 0)
  e2e    frames/s through the host-facing C-ABI call mmw_process_host(): pinned-host capture -> H2D ->
         chain -> D2H of the ordered detection list, every step
  roofline      dominant kernel: algorithmic bytes per launch / CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline  the plain-C oracle (oracle/mmw_oracle.c, kind "port": the reference has no CPU code for
                these stages) timed on the box's host cores on a bounded sample of the same workload
Workloads (BASELINE.json configs): cfg3 = 512 samples x 256 chirps x 12 virtual antennas (the configuration
north_star's >= 60 % roofline target is quoted on; default), cfg2 = 256 x 128 x 4, batch 1024.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

WORKLOADS = {
    # name: (S, C, A, frames per GPU per step, BASELINE.json config index)
    "cfg3": (512, 256, 12, 64, 2),
    "cfg2": (256, 128, 4, 1024, 1),
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
# `ncu --set full` capture under profiles/ (None until a capture for this exact build exists)
NCU_TRAFFIC_BYTES = {"cfg3": None, "cfg2": None}      # filled from profiles/ncu_r1_*.md


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


class CpuPort:
    """The plain-C oracle (fp64, radix-2 FFTs, frame-parallel with OpenMP) as a CPU baseline: a fixed pool of
    distinct synthetic frames of the workload, processed over and over until the requested amount of work is done."""

    def __init__(self, orc, pkg, S, C, A, cfg_id, cores):
        self.orc, self.S, self.C, self.A, self.cores = orc, S, C, A, cores
        self.pool = pkg.synth.cube_batch(max(cores, 8), S, C, A, cfg=cfg_id)
        self.wr, self.wd = orc.hann_periodic(S), orc.hann_periodic(C)
        self.run(1)                                                      # touch code and pages

    def run(self, passes):
        """processes `passes` x pool frames; returns (frames, seconds, detections of the last pass)"""
        n = self.pool.shape[0]
        t0 = time.perf_counter()
        for _ in range(passes):
            out = self.orc.process_frames(self.pool, n, self.S, self.C, self.A, self.wr, self.wd, n_threads=self.cores)
        return passes * n, time.perf_counter() - t0, int(out["n_total"])

    def passes_for(self, seconds):
        n, dt, _ = self.run(1)
        return max(1, int(round(seconds / max(dt, 1e-6))))


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The reference has no CPU (or GPU)
    code for the range/Doppler/CFAR/angle chain, so per north_star the plain-C oracle port
    stands in (kind "port"), frame-parallel on all host cores, on a bounded sample per step."""
    if rank != 0:
        return
    pkg, orc = entry.load_package(), entry.load_oracle()
    orc.build()
    S, C, A, _, cfg_idx = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    port = CpuPort(orc, pkg, S, C, A, cfg_idx + 1, cores)
    passes = port.passes_for(1.0)                                     # about one second of wall clock per step
    per_step = passes * port.pool.shape[0]
    times = []
    for i in range(args.warmup + args.steps):
        _, dt, _ = port.run(passes)
        if i >= args.warmup:
            times.append(dt)
    T = float(np.sum(times))
    value = per_step * len(times) / T
    line = {
        "impl": "reference", "metric": "radar frames/sec (ADC cube->detections)", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * T / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {S} samples x {C} chirps x {A} antennas, 2-D CA-CFAR, angle FFT "
                               f"(BASELINE.json configs[{cfg_idx}])", "frames_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{per_step} frames per step x {len(times)} steps, plain-C fp64 oracle, {cores} threads"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU per step (default: workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--keep-cube", action="store_true", help="materialise the Doppler cube in HBM (default: fused)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    # keep stdout for the one JSON line: libraries (NCCL's version banner) write to fd 1 during init
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    pkg = entry.load_package()
    S, C, A, F, cfg_idx = WORKLOADS[args.workload]
    if args.frames:
        F = args.frames
    K, W = args.steps, args.warmup

    ctx = pkg.RadarContext(S, C, A, F, keep_doppler_cube=args.keep_cube, max_det_per_frame=4096, device=local_rank)
    first_frame = rank * F                                     # weak scaling: every rank owns F frames of the global batch
    ctx.set_frame_offset(first_frame)
    adc = pkg.synth.cube_batch_torch(F, S, C, A, dev, cfg=cfg_idx + 1, first_frame=first_frame)
    torch.cuda.synchronize()
    stream = torch.cuda.Stream(device=dev)
    ctx.use_stream(stream.cuda_stream)
    dense_ptr, header_ptr = ctx.device_results()
    header_view = pkg.sharding.device_bytes_view(header_ptr, 16, dev)
    # exchange step (N > 1 only): fixed-size NCCL gather of each rank's result block to rank 0 + one merge kernel
    gather_records = min(F * ctx.max_det_per_frame, 32768)
    gather = pkg.sharding.DetectionGather(ctx, dev, gather_records) if world > 1 else None

    def step():
        ctx.process_device(adc, F)
        return gather.run() if gather is not None else None

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(W):
            step()
        sync_all()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(K):
            gathered = step()
        e1.record(stream)
        sync_all()
        ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    frame_counts = ctx.read_counts(F)                       # true per-frame hit counts of this rank's last batch
    if world == 1:
        n_det_step, gather_overflow = int(header_view[:4].view(torch.int32).item()), 0
    elif rank == 0:
        recs, hdr = gather.read(pkg.DET_DTYPE)
        n_det_step, gather_overflow = len(recs), int(hdr[3])
        assert np.all(np.diff(recs["frame"].astype(np.int64)) >= 0) and int(hdr[2]) == world * F     # ordered, all frames accounted for
    else:
        n_det_step, gather_overflow = 0, 0

    # ---- per-stage device times (events between launches) for the roofline of the dominant kernel ----
    ctx.use_stream(None)
    total_ms, stage_ms = ctx.time_device(adc, F, max(3, min(K, 10)), per_stage=True)
    iters = max(3, min(K, 10))
    stage_ms = [s / iters for s in stage_ms]
    N_adc, N, M = S * C * A, ctx.Sp * ctx.Cp * A, ctx.Sp * ctx.Cp
    stage_bytes = [
        F * (4 * N_adc + 8 * A * ctx.Sp * C),                   # K1: int16 IQ in, range spectrum out
        F * (8 * A * ctx.Sp * C + (8 * N if args.keep_cube else 0) + 4 * M),   # K2: spectrum in, (cube +) power map out
        F * (4 * M + M // 8),                                   # K3: power map in, bit mask out
        F * (M // 8),                                           # K4+K5: mask in (+ D records)
    ]
    names = ["range_fft_kernel", "doppler_fft_kernel", "cfar_kernel", "detect_kernel+compact_kernel"]
    dom = int(np.argmax(stage_ms))
    peak, peak_src = peaks()
    achieved = stage_bytes[dom] / (stage_ms[dom] * 1e-3) / 1e9
    b_alg = int(ctx.info.algorithmic_bytes_per_frame)
    fps = world * F * K / (ms * 1e-3)

    # ---- end to end through the host-facing API: pinned host capture -> H2D -> chain -> D2H list ----
    host = torch.empty((F, ctx.frame_shorts), dtype=torch.int16, pin_memory=True)
    host.copy_(adc)
    torch.cuda.synchronize()
    out = np.empty(F * ctx.max_det_per_frame, pkg.DET_DTYPE)
    e2e_steps = max(3, min(K, 10))
    for _ in range(2):
        dets, _ = ctx.process_host(host, F, out=out)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        dets, _ = ctx.process_host(host, F, out=out)
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_fps = world * F * e2e_steps / t_e2e
    clocks = sampler.stop() if rank == 0 else None      # sampled every 200 ms from the timed region to the end of the e2e loop

    if rank == 0:
        line = {
            "metric": "radar frames/sec (ADC cube->detections)", "value": fps, "unit": "frames/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"{args.workload}: {S} samples x {C} chirps x {A} antennas, 2-D CA-CFAR (guard 2x2, train 8x4, alpha 15), "
                            f"{ctx.n_theta}-pt angle FFT (BASELINE.json configs[{cfg_idx}])",
                "frames_per_gpu_per_step": F, "sharding": f"frame-sharded x{world}" + ("" if world == 1 else f"; per step one NCCL gather of a fixed {32 + 24 * gather_records}-byte result block per rank to rank 0 + merge kernel (overflow={gather_overflow})"),
                "doppler_cube": "materialised" if args.keep_cube else "fused (not written to HBM)",
                "l2": f"inputs larger than L2: {F * 4 * N_adc / 1e6:.0f} MB int16 capture + {F * 8 * A * ctx.Sp * C / 1e6:.0f} MB intermediate per step vs 126 MB L2",
                "detections_per_step": n_det_step, "max_detections_in_one_frame": int(frame_counts.max()),
                "max_det_per_frame": ctx.max_det_per_frame,
            },
            "roofline": {
                "bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": NCU_TRAFFIC_BYTES.get(args.workload), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": stage_bytes[dom], "kernel_ms": stage_ms[dom],
                "stage_ms": dict(zip(names, stage_ms)),
                "stage_gbs": {n: (b / (t * 1e-3) / 1e9 if t > 0 else None) for n, b, t in zip(names, stage_bytes, stage_ms)},
                "pipeline": {"algorithmic_bytes_per_frame": b_alg, "achieved": b_alg * F * K / (ms * 1e-3) / 1e9,
                             "frac": b_alg * F * K / (ms * 1e-3) / 1e9 / peak, "frac_of_8000_nominal": b_alg * F * K / (ms * 1e-3) / 1e9 / 8000.0,
                             "note": "28N+8M bytes/frame (SURVEY.md 8d) x frames / step time, per GPU"},
            },
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": F * 4 * N_adc,
                    "d2h_bytes_per_step": 16 + 24 * len(dets), "steps": e2e_steps, "api": "mmw_process_host"},
            "gpu_launches": K * (ctx.info.kernels_per_batch + (1 if world > 1 else 0)),
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            orc = entry.load_oracle()
            orc.build()
            cores = os.cpu_count() or 1
            port = CpuPort(orc, pkg, S, C, A, cfg_idx + 1, cores)
            passes = port.passes_for(12.0)                             # about 12 s of wall clock on all host cores
            n, dt, _ = port.run(passes)
            line["cpu_baseline"] = {"value": n / dt, "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": f"{n} frames ({port.pool.shape[0]} distinct, {passes} passes) of the same workload in {dt:.1f} s; "
                                              f"plain-C fp64 oracle (the reference has no CPU code for these stages), {cores} OpenMP threads"}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
