"""Throughput of the device-resident chain with one batch in flight vs two (two contexts, each on its own stream):
does the tail of one batch's persistent FFT kernels / its latency-bound detection kernels fill with the other batch's work?"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
SHAPES = {"cfg3": (512, 256, 12, 64), "cfg2": (256, 128, 4, 1024), "cfg5": (256, 128, 12, 64)}
dev = torch.device("cuda", 0)
CAPS = [int(c) for c in os.environ.get("PROBE_CAPS", "0").split(",")]      # MMW_CTAS_PER_SM values to try (0 = no cap)
for wl, cap in [(w, c) for w in (sys.argv[1:] or ["cfg3", "cfg2"]) for c in CAPS]:
    S, C, A, F = SHAPES[wl]
    os.environ["MMW_CTAS_PER_SM"] = str(cap)
    for depth in (1, 2, 3, 4):
        ctxs = [pkg.RadarContext(S, C, A, F) for _ in range(depth)]
        adcs = [pkg.synth.cube_batch_torch(F, S, C, A, dev, cfg=3, first_frame=i * F) for i in range(depth)]
        K = 30
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(K):
                ctxs[k % depth].process_device(adcs[k % depth], F)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        n = [len(c.read_detections()[0]) for c in ctxs]
        print(f"{wl} ctas/SM cap {cap} in flight {depth}: {dt / K * 1e3:.4f} ms per batch, {F * K / dt:.0f} frames/s, detections {n}", flush=True)
        for c in ctxs:
            c.close()
        del adcs
