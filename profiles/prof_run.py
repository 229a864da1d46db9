"""Small fixed run for ncu: 3 batches of a bench workload through the device-resident path.
    python profiles/prof_run.py [cfg3|cfg2|cfg4|cfg5] [scene cfg number, default: the bench's]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
S, C, A, F, scene, cap = {"cfg3": (512, 256, 12, 64, 3, 4096), "cfg2": (256, 128, 4, 1024, 2, 4096), "cfg4": (1024, 512, 192, 4, 4, 32768),
                          "cfg5": (256, 128, 12, 64, 5, 4096)}[wl]
scene = int(sys.argv[2]) if len(sys.argv) > 2 else scene
dev = torch.device("cuda", 0)
adc = pkg.synth.cube_batch_torch(F, S, C, A, dev, cfg=scene)
torch.cuda.synchronize()
with pkg.RadarContext(S, C, A, F, max_det_per_frame=cap) as ctx:
    for _ in range(3):
        ctx.process_device(adc, F)
    dets, ov = ctx.read_detections()
    print("detections", len(dets), "overflow", ov)
