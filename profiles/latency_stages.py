import os, sys
sys.path.insert(0, os.getcwd())
import torch
import __graft_entry__ as entry
pkg = entry.load_package()
S, C, A = 256, 128, 12
dev = torch.device("cuda", 0)
for F in (1, 2, 4, 8):
    adc = pkg.synth.cube_batch_torch(F, S, C, A, dev, cfg=5)
    with pkg.RadarContext(S, C, A, F) as ctx:
        ctx.time_device(adc, F, 5)
        tot, st = ctx.time_device(adc, F, 200, per_stage=True)
        n = 200
        print(f"F={F}: batch {tot/n*1e3:.1f} us | range {st[0]/n*1e3:.1f} doppler {st[1]/n*1e3:.1f} cfar {st[2]/n*1e3:.1f} detect {st[3]/n*1e3:.1f} us", flush=True)
