"""GPU: the rows SURVEY.md §8f lists next to the hot path — static-clutter removal, capture-file ingest, the
legacy file loop and the device-resident legacy entry point — each against the oracle or against the
already-verified in-memory path on the same bytes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(64, 64, 2), (100, 128, 4), (256, 128, 4), (512, 256, 12), (1024, 64, 2)])
def test_base_frame_subtraction_matches_oracle(pkg, orc, shape):
    S, C, A = shape
    F = 2
    adc = pkg.synth.cube_batch(F, S, C, A, cfg=6, n_targets=4)
    base = pkg.synth.cube(77, S, C, A, cfg=6, n_targets=2)          # static scene: two strong reflectors + its own noise
    adc = (adc.astype(np.int32) // 2 + base.astype(np.int32)[None, :] // 2).astype(np.int16)
    base = (base.astype(np.int32) // 2).astype(np.int16)
    with pkg.RadarContext(S, C, A, F, keep_doppler_cube=True) as ctx:
        wr, wd = ctx.get_windows()
        ctx.set_base_frame(base)
        dets, ov = ctx.process_host(adc, F)
        rs = ctx.range_spectrum(1)
        P = ctx.power_map(1)
        ref = orc.process_frames(adc, F, S, C, A, wr, wd, want=("rs", "P", "noise"), base=base, n_threads=4)
        rs_ref = ref["rs"][1] * wd[None, None, :]
        assert np.abs(rs - rs_ref).max() / np.abs(rs_ref).max() < 1e-4
        assert np.abs(P - ref["P"][1]).max() / ref["P"][1].max() < 1e-4
        thr = 15.0 * ref["noise"]
        near = np.abs(ref["P"] - thr) <= 1e-5 * thr
        got = {(int(d["frame"]), int(d["range_bin"]), int(d["doppler_bin"])) for d in dets}
        want = {(int(d["frame"]), int(d["range_bin"]), int(d["doppler_bin"])) for d in ref["dets"]}
        assert not {k for k in got ^ want if not near[k]} and not ov
        # turning it off again restores the plain chain bit for bit
        ctx.set_base_frame(None)
        a, _ = ctx.process_host(adc, F)
    with pkg.RadarContext(S, C, A, F, keep_doppler_cube=True) as ctx2:
        b, _ = ctx2.process_host(adc, F)
    assert a.tobytes() == b.tobytes()


def test_frame_equal_to_base_gives_nothing(pkg):
    S, C, A = 256, 128, 4
    adc = pkg.synth.cube_batch(2, S, C, A, cfg=2)
    with pkg.RadarContext(S, C, A, 2) as ctx:
        ctx.set_base_frame(adc[1])
        dets, _ = ctx.process_host(adc, 2)
        assert not ctx.power_map(1).any()
        assert len(dets) > 0 and not (dets["frame"] == 1).any()


def test_capture_file_equals_in_memory_path(pkg, tmp_path):
    S, C, A, B = 128, 64, 4, 4                     # batches of 4 frames, file of 11 frames: 3 batches, the last one ragged
    adc = pkg.synth.cube_batch(11, S, C, A, cfg=8, n_targets=3)
    path = tmp_path / "capture.bin"
    adc.tofile(path)
    with pkg.RadarContext(S, C, A, 16) as big:
        want, _ = big.process_host(adc, 11)
    with pkg.RadarContext(S, C, A, B) as ctx:
        got, n, ov = ctx.process_capture_file(str(path))
        assert n == 11 and not ov and got.tobytes() == want.tobytes()
        # window of the file: frames 3..7, numbered by their position in the file
        got, n, _ = ctx.process_capture_file(str(path), first_frame=3, max_frames=5)
        sel = want[(want["frame"] >= 3) & (want["frame"] < 8)]
        assert n == 5 and got.tobytes() == sel.tobytes()
        # past the end: nothing to do, no error
        got, n, _ = ctx.process_capture_file(str(path), first_frame=11)
        assert n == 0 and len(got) == 0
        # truncated file: the partial last frame is zero-filled and processed
        raw = adc.tobytes()[: 10 * adc.shape[1] * 2 + 1000]
        (tmp_path / "short.bin").write_bytes(raw)
        padded = np.frombuffer(raw + bytes(adc.shape[1] * 2 - 1000), np.int16).reshape(11, -1)
        got, n, _ = ctx.process_capture_file(str(tmp_path / "short.bin"))
        with pkg.RadarContext(S, C, A, 16) as big:
            want_short, _ = big.process_host(padded, 11)
        assert n == 11 and got.tobytes() == want_short.tobytes()
        # detection capacity smaller than the list: truncated, flagged, ordered prefix
        got, n, ov = ctx.process_capture_file(str(path), det_capacity=7)
        assert ov and len(got) == 7 and got.tobytes() == want[:7].tobytes()


def test_capture_file_with_first_frame_as_base(pkg, orc, tmp_path):
    """the reference's cudaTiming() convention: frame 0 of the file is the base frame and is not processed"""
    S, C, A = 100, 128, 4
    adc = pkg.synth.cube_batch(5, S, C, A, cfg=9, n_targets=3) // 2
    path = tmp_path / "fhy_like.bin"
    adc.tofile(path)
    with pkg.RadarContext(S, C, A, 2) as ctx:
        wr, wd = ctx.get_windows()
        got, n, _ = ctx.process_capture_file(str(path), use_first_as_base=True)
        ref = orc.process_frames(adc[1:], 4, S, C, A, wr, wd, want=("P", "noise"), base=adc[0], n_threads=4)
    assert n == 4
    thr = 15.0 * ref["noise"]
    near = np.abs(ref["P"] - thr) <= 1e-5 * thr
    g = {(int(d["frame"]) - 1, int(d["range_bin"]), int(d["doppler_bin"])) for d in got}      # file frame 1 = oracle frame 0
    w = {(int(d["frame"]), int(d["range_bin"]), int(d["doppler_bin"])) for d in ref["dets"]}
    assert not {k for k in g ^ w if not near[k]} and len(w) > 0


def test_legacy_file_loop_matches_reference_cpu_path(pkg, orc, tmp_path):
    cap = pkg.synth.legacy_capture(12, seed=4, moving=True)          # the fhy_s.bin stand-in
    path = tmp_path / "fhy_direct.bin"
    cap.tofile(path)
    dist, raw, n = pkg.api.legacy_process_file(str(path))
    assert n == 11
    base = orc.reshape(cap[0], 100, 128, 4)[:12800]
    for f in range(1, 12):
        d_ref, raw_ref = orc.legacy_frame(cap[f], base)
        assert raw[f - 1] == raw_ref and dist[f - 1] == d_ref
    # ragged tail: the short final frame is processed with its short count, like the reference loop
    (tmp_path / "ragged.bin").write_bytes(cap.tobytes()[: 3 * 204800 + 50000])
    dist2, raw2, n2 = pkg.api.legacy_process_file(str(tmp_path / "ragged.bin"))
    assert n2 == 3 and np.array_equal(raw2[:2], raw[:2])
    d_ref, raw_ref = orc.legacy_frame(cap[3][:25000], base)
    assert raw2[2] == raw_ref and dist2[2] == d_ref


def test_legacy_device_resident_entry(pkg, orc):
    import torch

    cap = pkg.synth.legacy_capture(9, seed=5)
    base = orc.reshape(cap[0], 100, 128, 4)[:12800]
    dev = torch.device("cuda", 0)
    frames = torch.from_numpy(cap[1:]).to(dev)
    raw = torch.empty(8, dtype=torch.int32, device=dev)
    pkg.api.legacy_process_device(frames, 8, base, raw)
    pkg.api.legacy_sync()
    want = [orc.legacy_frame(cap[f], base)[1] for f in range(1, 9)]
    assert raw.cpu().tolist() == want


def test_graph_mode_is_bit_identical_to_eager_launches(pkg):
    """cfg5's per-frame latency path: captured-graph replays give exactly the eager results, for changing inputs,
    frame offsets, base frames and both entry points."""
    import torch

    S, C, A, F = 256, 128, 12, 3
    adc = pkg.synth.cube_batch(F, S, C, A, cfg=5, n_targets=6)
    dev = torch.device("cuda", 0)
    adc_dev = torch.from_numpy(adc).to(dev)
    with pkg.RadarContext(S, C, A, F) as eager, pkg.RadarContext(S, C, A, F) as graphed:
        graphed.set_graph_mode(True)
        for rep in range(3):                                  # rep 0 warms (eager), 1 captures, 2 replays
            for f in range(F):
                for ctx in (eager, graphed):
                    ctx.set_frame_offset(10 * f)
                a, _ = eager.process_host(adc[f], 1)
                b, _ = graphed.process_host(adc[f], 1)
                assert a.tobytes() == b.tobytes() and len(a) > 0 and a["frame"][0] == 10 * f
                eager.process_device(adc_dev[f], 1)
                graphed.process_device(adc_dev[f], 1)
                assert eager.read_detections()[0].tobytes() == graphed.read_detections()[0].tobytes()
        for ctx in (eager, graphed):
            ctx.set_frame_offset(0)
            ctx.set_base_frame(adc[0])
        a, _ = eager.process_host(adc, F)
        b, _ = graphed.process_host(adc, F)
        assert a.tobytes() == b.tobytes() and not (a["frame"] == 0).any()
        b2, _ = graphed.process_host(adc, F)                  # replay of the batch graph
        assert b2.tobytes() == a.tobytes()
