"""Per-stage device times of the chain (CUDA events between launches, mmw_time_device) for the bench workloads.
Run on the GPU box:  python profiles/stage_times.py [cfg3 cfg2 cfg4 ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
SHAPES = {"cfg3": (512, 256, 12, 64), "cfg2": (256, 128, 4, 1024), "cfg4": (1024, 512, 192, 4), "cfg5": (256, 128, 12, 64),
          "legacy": (100, 128, 4, 256)}
dev = torch.device("cuda", 0)
with_base = "--base" in sys.argv          # static-clutter removal on: K1 also reads the base frame (from L2)
for wl in ([a for a in sys.argv[1:] if not a.startswith("--")] or ["cfg3", "cfg2"]):
    S, C, A, F = SHAPES[wl]
    adc = pkg.synth.cube_batch_torch(F, S, C, A, dev, cfg=3)
    for keep in (False, True):
        with pkg.RadarContext(S, C, A, F, keep_doppler_cube=keep) as ctx:
            if with_base:
                ctx.set_base_frame(adc[0].cpu().numpy())
            ctx.time_device(adc, F, 3)
            tot, st = ctx.time_device(adc, F, 20, per_stage=True)
            n = 20
            balg = ctx.info.algorithmic_bytes_per_frame * F
            print(f"{wl} {S}x{C}x{A} F={F} keep={int(keep)}{' base' if with_base else ''}: total {tot / n:.4f} ms = {F / (tot / n) * 1e3:.0f} frames/s, "
                  f"{balg / (tot / n) / 1e6:.0f} GB/s of B_alg | range {st[0] / n:.4f} doppler {st[1] / n:.4f} cfar {st[2] / n:.4f} detect {st[3] / n:.4f}",
                  flush=True)
    del adc
