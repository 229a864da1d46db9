"""K2 frame order: MMW_K2_REVERSE = 1 lets the Doppler kernel walk a batch from its last frame to its first, so that it starts on
the range spectrum K1 wrote last (still in the 126 MB L2).  Stage times; the detections must not change."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
SHAPES = {"cfg3": (512, 256, 12, 64), "cfg2": (256, 128, 4, 1024), "cfg5": (256, 128, 12, 64), "cfg4": (1024, 512, 192, 4)}
dev = torch.device("cuda", 0)
for wl in (sys.argv[1:] or ["cfg3", "cfg2", "cfg4", "cfg5"]):
    S, C, A, F = SHAPES[wl]
    adc = pkg.synth.cube_batch_torch(F, S, C, A, dev, cfg=3)
    with pkg.RadarContext(S, C, A, F, max_det_per_frame=4096) as ctx:
        ref = None
        for var in (0, 1, 0, 1):
            os.environ["MMW_K2_REVERSE"] = str(var)
            ctx.process_device(adc, F)
            dets, _ = ctx.read_detections()
            ref = dets.tobytes() if ref is None else ref
            ctx.time_device(adc, F, 3)
            tot, st = ctx.time_device(adc, F, 20, per_stage=True)
            tot2 = ctx.time_device(adc, F, 20)
            print(f"{wl} reverse={var}: doppler {st[1] / 20:.4f} ms, range {st[0] / 20:.4f} (instrumented total {tot / 20:.4f}, back-to-back total {tot2 / 20:.4f}) | same bytes: {dets.tobytes() == ref}", flush=True)
    os.environ.pop("MMW_K2_REVERSE", None)
    del adc
