"""Small fixed run for ncu: 2 batches of the cfg3 shape (8 frames) through the device-resident path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
S, C, A, F = {"cfg3": (512, 256, 12, 64), "cfg2": (256, 128, 4, 1024)}[wl]
dev = torch.device("cuda", 0)
adc = pkg.synth.cube_batch_torch(F, S, C, A, dev, cfg=3)
torch.cuda.synchronize()
with pkg.RadarContext(S, C, A, F) as ctx:
    for _ in range(3):
        ctx.process_device(adc, F)
    dets, ov = ctx.read_detections()
    print("detections", len(dets), "overflow", ov)
