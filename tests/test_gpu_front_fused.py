"""Fused front (MMW_FRONT=2/3: range FFT and Doppler FFT as two roles of one cooperative kernel, csrc/mmw_front.cuh): the
same bits as the two-kernel chain, and the oracle's detections.  The switch is read at mmw_create."""
import os

import numpy as np
import pytest

import __graft_entry__ as entry

pytestmark = pytest.mark.gpu


def _run(pkg, adc, S, C, A, F, mode):
    old = {k: os.environ.get(k) for k in ("MMW_FRONT", "MMW_FRONT_STATS")}
    os.environ["MMW_FRONT"] = str(mode)
    os.environ["MMW_FRONT_STATS"] = "1"
    try:
        with pkg.RadarContext(S, C, A, F, max_det_per_frame=4096) as ctx:
            dets, overflow = ctx.process_host(adc, F)
            maps = np.stack([ctx.power_map(f) for f in range(F)])
            wr, wd = ctx.get_windows()
            ran_fused = bool(ctx.front_stats()[:, 2].any())          # the fused kernel leaves its per-CTA record
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert not overflow
    assert ran_fused == (mode >= 2), f"MMW_FRONT={mode}: fused front kernel {'did not run' if mode >= 2 else 'ran'}"
    return dets, maps, wr, wd


# (S, C, A, F): the benchmark shape, a padded one (S < Sp, C < Cp), a frame count that is no multiple of the consumer group,
# and the 256 x 128 plan
@pytest.mark.parametrize("shape", [(512, 256, 12, 16), (500, 250, 3, 30), (512, 256, 2, 37), (256, 128, 12, 40), (200, 100, 4, 90)])
@pytest.mark.parametrize("mode", [2, 3])
def test_fused_front_gives_the_bits_of_the_two_kernel_chain(shape, mode):
    S, C, A, F = shape
    if mode == 3 and not (512 >= S > 256):
        pytest.skip("the 4-warp form exists for the 512 x 256 plan only")
    pkg = entry.load_package()
    adc = pkg.synth.cube_batch(F, S, C, A, cfg=3, n_targets=4)
    d1, m1, _, _ = _run(pkg, adc, S, C, A, F, 1)
    d2, m2, _, _ = _run(pkg, adc, S, C, A, F, mode)
    assert np.array_equal(m1, m2), "power maps differ between the fused front and the two-kernel chain"
    assert d1.tobytes() == d2.tobytes(), "detection records differ"


def test_fused_front_against_the_oracle():
    pkg = entry.load_package()
    orc = entry.load_oracle()
    S, C, A, F = 512, 256, 4, 20
    adc = pkg.synth.cube_batch(F, S, C, A, cfg=3, n_targets=4)
    dets, maps, wr, wd = _run(pkg, adc, S, C, A, F, 2)
    ref = orc.process_frames(adc, F, S, C, A, wr, wd, want=("P", "noise"))
    for f in range(F):
        err = np.abs(maps[f] - ref["P"][f]).max() / ref["P"][f].max()
        assert err < 1e-4, f"frame {f}: power map off by {err:.2e} of its maximum"       # tolerance: north_star's fp32 1e-4
    thr = 15.0 * ref["noise"]
    near = np.abs(ref["P"] - thr) <= 1e-5 * thr
    got = {(int(d["frame"]), int(d["range_bin"]), int(d["doppler_bin"])) for d in dets}
    want = {(int(d["frame"]), int(d["range_bin"]), int(d["doppler_bin"])) for d in ref["dets"]}
    assert not {k for k in got ^ want if not near[k]}, "detections differ from the oracle away from the threshold"
