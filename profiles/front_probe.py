"""Fused front (MMW_FRONT=2: K1 and K2 as two roles of one cooperative kernel, range spectrum read back out of the L2)
against the two-kernel chain (MMW_FRONT=1): same bits required, device times compared.
    python profiles/front_probe.py [cfg3 cfg5 cfg2] [--windows 0,16,32,64]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
SHAPES = {"cfg3": (512, 256, 12, 64), "cfg2": (256, 128, 4, 1024), "cfg5": (256, 128, 12, 64), "pad": (500, 250, 12, 32),
          "odd": (512, 256, 12, 37)}
dev = torch.device("cuda", 0)
windows = [0]
names = []
for a in sys.argv[1:]:
    if a.startswith("--windows"):
        windows = [int(x) for x in a.split("=")[1].split(",")]
    else:
        names.append(a)
for wl in names or ["cfg3"]:
    S, C, A, F = SHAPES[wl]
    adc = pkg.synth.cube_batch_torch(F, S, C, A, dev, cfg=3)
    ref = None
    modes = [int(x) for x in os.environ.get("FRONT_MODES", "2").split(",")]
    for mode, win in [(1, 0)] + [(m, w) for m in modes for w in windows]:
        os.environ["MMW_FRONT"] = str(mode)
        os.environ["MMW_FRONT_WINDOW"] = str(win)
        with pkg.RadarContext(S, C, A, F, max_det_per_frame=4096) as ctx:
            ctx.time_device(adc, F, 3)
            tot, st = ctx.time_device(adc, F, 20, per_stage=True)
            ctx.process_device(adc, F)
            dets, ov = ctx.read_detections()
            pm = np.stack([ctx.power_map(f) for f in (0, F // 2, F - 1)])
            if ref is None:
                ref = (pm, dets.tobytes())
                same = "reference"
            else:
                same = f"pmap identical {np.array_equal(pm, ref[0])}, detections identical {dets.tobytes() == ref[1]}"
            if mode >= 2 and os.environ.get("MMW_FRONT_STATS"):
                st_ = ctx.front_stats()
                st_ = st_[st_[:, 2] > 0]
                role = (st_[:, 0] >> np.uint64(32)).astype(int)
                smid = (st_[:, 0] & np.uint64(0xffffffff)).astype(int)
                t0 = st_[:, 1].min()
                for r, name in ((1, "range role"), (0, "doppler role")):
                    m = role == r
                    if not m.any():
                        continue
                    dur = (st_[m, 2] - st_[m, 1]).astype(float) / 1e3
                    end = (st_[m, 2] - t0).astype(float) / 1e3
                    wait = st_[m, 3].astype(float) / 1e3
                    print(f"    {name}: {m.sum()} CTAs on {len(set(smid[m]))} SMs; busy {dur.mean():.1f} us (min {dur.min():.1f}, max {dur.max():.1f}); "
                          f"ends at {end.mean():.1f} us (max {end.max():.1f}); waited on the other role {wait.mean():.1f} us (max {wait.max():.1f}); units {st_[m, 4].mean():.1f}")
                both = len(set(smid[role == 1]) & set(smid[role == 0]))
                cnt = np.bincount(smid, minlength=148)
                per = [((role[smid == s_] == 1).sum(), (role[smid == s_] == 0).sum()) for s_ in sorted(set(smid))]
                from collections import Counter
                print(f"    (range, doppler) CTAs per SM: {dict(Counter(per))}")
                print(f"    SMs hosting both roles: {both}; CTAs per SM: min {cnt.min()} max {cnt.max()}; SMs with 2 range CTAs: {sum(1 for s_ in set(smid) if (role[smid == s_] == 1).sum() == 2)}; with 2 doppler CTAs: {sum(1 for s_ in set(smid) if (role[smid == s_] == 0).sum() == 2)}")
            print(f"{wl} {S}x{C}x{A} F={F} front={mode} window={win}: total {tot / 20:.4f} ms = {F / (tot / 20) * 1e3:.0f} frames/s | "
                  f"front {(st[0] + st[1]) / 20:.4f} cfar {st[2] / 20:.4f} detect {st[3] / 20:.4f} | {len(dets)} detections | {same}", flush=True)
    del adc
