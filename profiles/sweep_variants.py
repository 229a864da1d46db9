"""Sweeps MMW_K1_VARIANT / MMW_K2_VARIANT (tile-shape experiments compiled into the library) and prints stage times."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
SHAPES = {"cfg3": (512, 256, 12, 64), "cfg2": (256, 128, 4, 1024), "cfg4": (1024, 512, 192, 4)}
dev = torch.device("cuda", 0)
nv1, nv2 = int(os.environ.get("NV1", 1)), int(os.environ.get("NV2", 6))
for wl in (sys.argv[1:] or ["cfg3", "cfg2"]):
    S, C, A, F = SHAPES[wl]
    adc = pkg.synth.cube_batch_torch(F, S, C, A, dev, cfg=3)
    with pkg.RadarContext(S, C, A, F) as ctx:
        for v1 in range(nv1):
            for v2 in range(nv2):
                if v1 and v2:
                    continue
                os.environ["MMW_K1_VARIANT"], os.environ["MMW_K2_VARIANT"] = str(v1), str(v2)
                ctx.time_device(adc, F, 3)
                tot, st = ctx.time_device(adc, F, 20, per_stage=True)
                print(f"{wl} k1v={v1} k2v={v2}: total {tot / 20:.4f} ms | range {st[0] / 20:.4f} doppler {st[1] / 20:.4f} cfar {st[2] / 20:.4f} detect {st[3] / 20:.4f}", flush=True)
    del adc
