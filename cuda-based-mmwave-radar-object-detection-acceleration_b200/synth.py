"""Deterministic synthetic ADC captures in the reference's on-disk format.

Packing and frame order follow the reference reader (cudaBenchMarking.cpp:156-165,
:168-180): little-endian int16, per frame [chirp][antenna][sample], samples in
groups of four shorts [I(2m) I(2m+1) Q(2m) Q(2m+1)].  The bundled captures
(fhy_direct.bin / fhy_s.bin) are absent from the reference mount
(.MISSING_LARGE_BLOBS), so these stand in for them.
"""
from __future__ import annotations

import numpy as np

LEGACY_S, LEGACY_C, LEGACY_A = 100, 128, 4      # cudaBenchMarking.cpp:3-6
LEGACY_SHORTS = LEGACY_S * LEGACY_C * LEGACY_A * 2


def pack_iiqq(z: np.ndarray) -> np.ndarray:
    """complex [..., S] (S even) -> int16 [..., 2*S] in IIQQ groups."""
    S = z.shape[-1]
    assert S % 2 == 0
    re = np.rint(z.real).clip(-32768, 32767).astype(np.int16).reshape(*z.shape[:-1], S // 2, 2)
    im = np.rint(z.imag).clip(-32768, 32767).astype(np.int16).reshape(*z.shape[:-1], S // 2, 2)
    return np.concatenate([re, im], axis=-1).reshape(*z.shape[:-1], 2 * S)


def unpack_iiqq(s: np.ndarray) -> np.ndarray:
    """inverse of pack_iiqq -> complex128 [..., S]."""
    g = s.reshape(*s.shape[:-1], -1, 4).astype(np.float64)
    z = g[..., 0:2] + 1j * g[..., 2:4]
    return z.reshape(*s.shape[:-1], -1)


def legacy_capture(n_frames: int = 90, seed: int = 0, tone: float = 0.123, moving: bool = False) -> np.ndarray:
    """d1: reference-format capture, [n_frames][102400] int16.  Frame 0 is the
    base frame (noise only); later frames add one complex tone at `tone`
    cycles/sample on every chirp and rx (moving=True drifts it per frame: the
    fhy_s.bin stand-in)."""
    rng = np.random.default_rng(seed)
    S, C, A = LEGACY_S, LEGACY_C, LEGACY_A
    out = np.empty((n_frames, LEGACY_SHORTS), np.int16)
    k = np.arange(S)
    for f in range(n_frames):
        z = rng.normal(0, 20, (C, A, S)) + 1j * rng.normal(0, 20, (C, A, S))
        if f > 0:
            fr = tone + (0.0007 * f if moving else 0.0)
            z = z + 800.0 * np.exp(2j * np.pi * fr * k)[None, None, :]
        out[f] = pack_iiqq(z).reshape(-1)
    return out


def cube(frame_index: int, S: int, C: int, A: int, cfg: int = 0, n_targets: int = 8,
         noise_sigma: float = 30.0) -> np.ndarray:
    """d2..d5: one north-star frame, int16 [2*S*C*A].  Point targets on and
    between bins + complex Gaussian noise, saturated to +-16384.  The content
    depends only on (cfg, frame_index), never on which GPU processes it."""
    rng = np.random.default_rng(1000 * cfg + frame_index)
    s = np.arange(S)[None, None, :]
    c = np.arange(C)[:, None, None]
    a = np.arange(A)[None, :, None]
    z = rng.normal(0, noise_sigma, (C, A, S)) + 1j * rng.normal(0, noise_sigma, (C, A, S))
    for t in range(n_targets):
        f_r = rng.uniform(0.02, 0.98)
        f_d = rng.uniform(-0.5, 0.5)
        f_a = rng.uniform(-0.9, 0.9)
        if t % 2 == 0:                       # every other target exactly on a bin
            f_r = np.round(f_r * S) / S
            f_d = np.round(f_d * C) / C
        amp = rng.uniform(200, 4000)
        ph = rng.uniform(0, 2 * np.pi)
        z = z + amp * np.exp(1j * (2 * np.pi * (f_r * s + f_d * c + 0.5 * f_a * a) + ph))
    z = np.clip(z.real, -16384, 16384) + 1j * np.clip(z.imag, -16384, 16384)
    return pack_iiqq(z).reshape(-1)


def cube_batch(n_frames: int, S: int, C: int, A: int, cfg: int = 0, first_frame: int = 0, **kw) -> np.ndarray:
    return np.stack([cube(first_frame + f, S, C, A, cfg, **kw) for f in range(n_frames)])


def cube_batch_torch(n_frames: int, S: int, C: int, A: int, device, cfg: int = 0, first_frame: int = 0,
                     n_targets: int = 8, noise_sigma: float = 30.0, chunk: int = 8):
    """Same recipe generated on the device with torch (bench-size batches would
    take minutes in numpy).  Deterministic per (cfg, frame index) through a
    per-frame torch.Generator seed; not bit-identical to cube() (different RNG)."""
    import torch

    out = torch.empty((n_frames, 2 * S * C * A), dtype=torch.int16, device=device)
    s = torch.arange(S, device=device, dtype=torch.float32)[None, None, :]
    c = torch.arange(C, device=device, dtype=torch.float32)[:, None, None]
    a = torch.arange(A, device=device, dtype=torch.float32)[None, :, None]
    two_pi = 2.0 * np.pi
    for f in range(n_frames):
        g = torch.Generator(device=device)
        g.manual_seed(1000003 * cfg + first_frame + f)
        host = np.random.default_rng(1000 * cfg + first_frame + f)
        re = torch.randn((C, A, S), generator=g, device=device) * noise_sigma
        im = torch.randn((C, A, S), generator=g, device=device) * noise_sigma
        for t in range(n_targets):
            f_r = host.uniform(0.02, 0.98)
            f_d = host.uniform(-0.5, 0.5)
            f_a = host.uniform(-0.9, 0.9)
            if t % 2 == 0:
                f_r = np.round(f_r * S) / S
                f_d = np.round(f_d * C) / C
            amp = host.uniform(200, 4000)
            ph = host.uniform(0, 2 * np.pi)
            # reduce the phase mod 1 cycle in fp64-free form: fractional frequencies times small ints stay exact enough in fp32
            phase = torch.remainder(f_r * s + f_d * c + 0.5 * f_a * a, 1.0) * two_pi + ph
            re += amp * torch.cos(phase)
            im += amp * torch.sin(phase)
        re = re.clamp_(-16384, 16384).round_().to(torch.int16).reshape(C, A, S // 2, 2)
        im = im.clamp_(-16384, 16384).round_().to(torch.int16).reshape(C, A, S // 2, 2)
        out[f] = torch.cat([re, im], dim=-1).reshape(-1)
    return out
