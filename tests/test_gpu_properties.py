"""Size-independent properties at the full BASELINE.json shapes (the oracle is too slow to check every
frame there): exact power-of-two scaling, energy conservation, determinism, batch independence and
sharded == unsharded."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# (S, C, A, frames): cfg2, cfg3, cfg5's per-sensor cube and one frame of the cfg4 imaging cube (1024 x 512 x 192 = 403 MB)
FULL = [(256, 128, 4, 32), (512, 256, 12, 4), (256, 128, 12, 8), (1024, 512, 192, 1)]


@pytest.fixture(scope="module")
def batches(pkg):
    import torch

    out = {}
    for (S, C, A, F) in FULL:
        if S * C * A > 8 << 20:       # cfg4: generate on the device (numpy would take minutes), same recipe
            out[(S, C, A)] = pkg.synth.cube_batch_torch(F, S, C, A, torch.device("cuda", 0), cfg=4, n_targets=8).cpu().numpy()
            torch.cuda.empty_cache()
        else:
            out[(S, C, A)] = pkg.synth.cube_batch(F, S, C, A, cfg=2 if A == 4 else 3, n_targets=8)
    return out


@pytest.mark.parametrize("S,C,A,F", FULL)
def test_power_of_two_scaling_is_exact(pkg, batches, S, C, A, F):
    adc = batches[(S, C, A)]
    half = (adc // 2).astype(np.int16)                          # any int16 data; doubled stays in range (|x| <= 16384)
    with pkg.RadarContext(S, C, A, F) as ctx:
        d1, _ = ctx.process_host(half, F)
        p1 = ctx.power_map(F - 1)
        d2, _ = ctx.process_host((half * 2).astype(np.int16), F)
        p2 = ctx.power_map(F - 1)
    assert np.array_equal(p2, 4.0 * p1)                         # every fp32 op scales exactly by 2
    assert np.array_equal(d1["range_bin"], d2["range_bin"]) and np.array_equal(d1["doppler_bin"], d2["doppler_bin"])
    assert np.array_equal(d2["power"], 4 * d1["power"]) and np.array_equal(d1["angle_bin"], d2["angle_bin"])


@pytest.mark.parametrize("S,C,A,F", FULL)
def test_energy_conservation(pkg, batches, S, C, A, F):
    adc = batches[(S, C, A)]
    with pkg.RadarContext(S, C, A, F) as ctx:
        wr, wd = ctx.get_windows()
        ctx.process_host(adc, F)
        for f in sorted({0, F - 1}):
            e_in = 0.0
            for c0 in range(0, C, 32):                            # chunks of chirps keep the fp64 temporaries small
                z = pkg.synth.unpack_iiqq(adc[f].reshape(C, A, 2 * S)[c0:c0 + 32])
                e_in += (np.abs(z * wr[None, None, :] * wd[c0:c0 + 32, None, None]) ** 2).sum()
            e_out = ctx.power_map(f).astype(np.float64).sum()
            assert abs(e_out / (ctx.Sp * ctx.Cp) - e_in) <= 1e-5 * e_in       # Parseval, unnormalised 2-D FFT


@pytest.mark.parametrize("S,C,A,F", FULL)
def test_deterministic_and_batch_independent(pkg, batches, S, C, A, F):
    adc = batches[(S, C, A)]
    with pkg.RadarContext(S, C, A, F) as ctx:
        a, _ = ctx.process_host(adc, F)
        b, _ = ctx.process_host(adc, F)
        assert a.tobytes() == b.tobytes() and len(a) > 0
        keys = a["frame"].astype(np.int64) << 32 | a["range_bin"].astype(np.int64) << 16 | a["doppler_bin"]
        assert np.all(np.diff(keys) > 0)
        # a frame processed alone gives the records it gets inside the batch
        ctx.set_frame_offset(F - 1)
        one, _ = ctx.process_host(adc[F - 1:], 1)
        assert one.tobytes() == a[a["frame"] == F - 1].tobytes()


@pytest.mark.parametrize("S,C,A,F", FULL)
def test_sharded_equals_unsharded(pkg, batches, S, C, A, F):
    """two 'ranks' emulated as two contexts on one GPU, each with its own contiguous frame block"""
    adc = batches[(S, C, A)]
    with pkg.RadarContext(S, C, A, F) as ctx:
        whole, _ = ctx.process_host(adc, F)
    parts = []
    for rank in range(3):
        first, cnt = pkg.sharding.shard_frames(F, 3, rank)
        with pkg.RadarContext(S, C, A, max(cnt, 1)) as ctx:
            ctx.set_frame_offset(first)
            if cnt:
                parts.append(ctx.process_host(adc[first:first + cnt], cnt)[0])
    assert np.concatenate(parts).tobytes() == whole.tobytes()


def test_device_resident_path_and_stream(pkg, batches):
    import torch

    S, C, A, F = FULL[0]
    adc = batches[(S, C, A)]
    with pkg.RadarContext(S, C, A, F) as ctx:
        host, _ = ctx.process_host(adc, F)
        dev = torch.from_numpy(adc).cuda()
        s = torch.cuda.Stream()
        ctx.use_stream(s.cuda_stream)
        with torch.cuda.stream(s):
            ctx.process_device(dev, F)
        s.synchronize()
        got, _ = ctx.read_detections()
        assert got.tobytes() == host.tobytes()
        dense, header = ctx.device_results()
        hdr = pkg.sharding.device_bytes_view(header, 16, dev.device).cpu().numpy().view(np.uint32)
        assert hdr[0] == len(host) and hdr[2] == F and hdr[3] == 0
        recs = pkg.sharding.records_from_bytes(pkg.sharding.device_bytes_view(dense, 24 * len(host), dev.device), pkg.DET_DTYPE)
        assert recs.tobytes() == host.tobytes()
        ctx.use_stream(None)
        ms = ctx.time_device(dev, F, 2)
        assert ms > 0


def test_merge_of_gathered_rank_blocks_equals_unsharded(pkg, batches):
    """the rank-0 merge kernel (mmw_merge_gathered) on blocks produced by 3 emulated ranks == the 1-rank list"""
    import torch

    S, C, A, F = FULL[0]
    adc = batches[(S, C, A)]
    dev = torch.device("cuda", 0)
    with pkg.RadarContext(S, C, A, F) as ctx:
        whole, _ = ctx.process_host(adc, F)
    world, per_rank = 3, 4096
    stride = 32 + 24 * per_rank
    gathered = torch.zeros((world, stride), dtype=torch.uint8, device=dev)
    ctxs = []
    for rank in range(world):
        first, cnt = pkg.sharding.shard_frames(F, world, rank)
        c = pkg.RadarContext(S, C, A, cnt)
        c.set_frame_offset(first)
        c.process_device(torch.from_numpy(adc[first:first + cnt]).cuda(), cnt)
        c.read_detections()                                       # synchronise
        block, cap = c.device_result_block()
        gathered[rank].copy_(pkg.sharding.device_bytes_view(block, stride, dev))
        ctxs.append(c)
    merged = torch.zeros(32 + 24 * world * per_rank, dtype=torch.uint8, device=dev)
    ctxs[0].merge_gathered(gathered, world, stride, merged, world * per_rank)
    torch.cuda.synchronize()
    m = merged.cpu().numpy()
    hdr = m[:32].view(np.uint32)
    assert hdr[0] == len(whole) and hdr[2] == F and hdr[3] == 0
    assert m[32:32 + 24 * len(whole)].tobytes() == whole.tobytes()
    # a merged buffer that is too small truncates in order and raises the overflow word
    small = torch.zeros(32 + 24 * 100, dtype=torch.uint8, device=dev)
    ctxs[0].merge_gathered(gathered, world, stride, small, 100)
    torch.cuda.synchronize()
    s = small.cpu().numpy()
    assert s[:32].view(np.uint32)[0] == 100 and s[:32].view(np.uint32)[3] == 1 and s[32:].tobytes() == whole[:100].tobytes()
    for c in ctxs:
        c.close()


def test_contexts_on_two_devices_in_one_process(pkg, batches):
    """mmw_config.device: one process may drive several GPUs (kernel attributes and occupancy are cached per device)"""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    S, C, A, F = FULL[0]
    adc = batches[(S, C, A)]
    with pkg.RadarContext(S, C, A, F, device=1) as c1:         # device 1 FIRST: nothing has been configured on it yet
        d1, _ = c1.process_host(adc, F)
        with pkg.RadarContext(S, C, A, F, device=0) as c0:
            d0, _ = c0.process_host(adc, F)
            half = F // 2
            c1.set_frame_offset(half)
            a, _ = c0.process_host(adc[:half], half)
            b, _ = c1.process_host(adc[half:], F - half)
    assert len(d0) > 0 and d0.tobytes() == d1.tobytes()
    assert np.concatenate([a, b]).tobytes() == d0.tobytes()


def test_process_device_refuses_host_pointers(pkg, batches):
    S, C, A, F = FULL[0]
    adc = batches[(S, C, A)]
    with pkg.RadarContext(S, C, A, F) as ctx:
        with pytest.raises(pkg.RadarError, match="not a device pointer"):
            ctx.process_device(int(adc.ctypes.data) & ~15, F)
        good, _ = ctx.process_host(adc, F)                       # the context is still usable afterwards
        assert len(good) > 0


@pytest.mark.parametrize("graph", [False, True])
def test_submit_wait_two_batches_in_flight(pkg, batches, graph):
    """mmw_submit_host / mmw_wait: frames streamed through two contexts (two in flight) give, frame by frame, the bytes
    the synchronous mmw_process_host gives; a second submit or a wait with nothing queued is MMW_ERR_STATE."""
    import torch

    S, C, A, F = FULL[2]
    adc = batches[(S, C, A)]
    pinned = torch.empty(adc.shape, dtype=torch.int16, pin_memory=True)
    pinned.copy_(torch.from_numpy(adc))
    with pkg.RadarContext(S, C, A, 1) as sync_ctx:
        sync_ctx.set_graph_mode(graph)
        want = []
        for f in range(F):
            sync_ctx.set_frame_offset(f)
            want.append(sync_ctx.process_host(pinned[f], 1)[0].copy())
    ring = [pkg.RadarContext(S, C, A, 1) for _ in range(2)]
    try:
        for c in ring:
            c.set_graph_mode(graph)
        with pytest.raises(pkg.RadarError) as ei:
            ring[0].wait()
        assert ei.value.code == pkg.api.MMW_ERR_STATE
        got = [None] * F
        for f in range(F + 2):
            c = ring[f % 2]
            if f >= 2:
                got[f - 2] = c.wait()[0].copy()
            if f < F:
                c.set_frame_offset(f)
                c.submit_host(pinned[f], 1)
                if f == 0:
                    for call in (lambda: c.submit_host(pinned[f], 1), lambda: c.read_detections(),
                                 lambda: c.process_device(torch.zeros(c.frame_shorts, dtype=torch.int16, device="cuda"), 1)):
                        with pytest.raises(pkg.RadarError) as ei:
                            call()
                        assert ei.value.code == pkg.api.MMW_ERR_STATE
    finally:
        for c in ring:
            c.close()
    assert sum(len(w) for w in want) > 0
    for f in range(F):
        assert got[f].tobytes() == want[f].tobytes(), f
