"""Sweeps MMW_K3_VARIANT (CFAR kernel forms compiled into the library): 1 = tiled kernel, 2/3/4 = walk kernel with 64- /
128- / 256-bin strips, + 10 * nchunk = Doppler segment of 16 * nchunk bins; 0 = the launcher's own choice.  Prints stage
times and checks every variant's detections against the tiled kernel's (same cells; noise within 1e-5)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
SHAPES = {"cfg3": (512, 256, 12, 64), "cfg2": (256, 128, 4, 1024), "cfg4": (1024, 512, 192, 4), "cfg5": (256, 128, 12, 64),
          "cfg5x1": (256, 128, 12, 1)}
dev = torch.device("cuda", 0)
for wl in (sys.argv[1:] or ["cfg3", "cfg2", "cfg5", "cfg5x1"]):
    S, C, A, F = SHAPES[wl]
    adc = pkg.synth.cube_batch_torch(F, S, C, A, dev, cfg=3)
    with pkg.RadarContext(S, C, A, F, max_det_per_frame=4096) as ctx:
        ref = None
        for var in [1, 0, 22, 42, 82, 23, 43, 83, 24, 44, 84]:
            os.environ["MMW_K3_VARIANT"] = str(var)
            ctx.process_device(adc, F)
            dets, _ = ctx.read_detections()
            if ref is None:
                ref = dets.copy()
            same_cells = len(dets) == len(ref) and all(np.array_equal(dets[k], ref[k]) for k in ("frame", "range_bin", "doppler_bin"))
            noise_err = float(np.max(np.abs(dets["noise"] - ref["noise"]) / ref["noise"])) if same_cells and len(ref) else -1.0
            ctx.time_device(adc, F, 3)
            tot, st = ctx.time_device(adc, F, 20, per_stage=True)
            print(f"{wl} k3v={var}: cfar {st[2] / 20:.4f} ms (total {tot / 20:.4f}) | {len(dets)} detections, same cells as tiled: {same_cells}, "
                  f"max rel noise diff {noise_err:.2e}", flush=True)
        os.environ.pop("MMW_K3_VARIANT", None)
    del adc
