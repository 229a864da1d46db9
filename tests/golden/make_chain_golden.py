"""Generates tests/golden/chain_numpy.npz: the north-star stages (window, range FFT, Doppler FFT, power map, 2-D CA-CFAR,
angle arg-max, 3x3 grouping) computed with numpy alone — numpy.fft and brute-force window sums in fp64, no oracle code
— on seeded synthetic cubes of the reference's own frame shape (100 samples x 128 chirps x 4 rx), of 64 x 64 x 2, of a
192-antenna array (64 x 64 x 192) and of a ragged shape (68 x 66 x 3).

The reference has no code, tests or vectors for these stages ("parity unpinned"); this fixture pins the
definitions of DESIGN.md §2 independently of oracle/mmw_oracle.c, so that the oracle (tests/test_oracle_pin.py, no GPU)
and the CUDA chain (tests/test_gpu_parity.py) are both checked against numbers neither of them produced.

    python tests/golden/make_chain_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
GR, GD, TR, TD, ALPHA = 2, 2, 8, 4, 15.0
# S, C, A, frames, synth cfg: the reference's frame shape, a small square cube, a wide array (A > 64: the 256-point angle FFT of
# the imaging configuration) and a ragged shape zero-padded on both axes (68 -> 128 samples, 66 -> 128 chirps)
CASES = [(100, 128, 4, 2, 21), (64, 64, 2, 2, 22), (64, 64, 192, 1, 23), (68, 66, 3, 1, 24)]


def next_pow2(n):
    return 1 << (n - 1).bit_length()


def hann(n):
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)).astype(np.float32)      # periodic, fp64 -> fp32 table


def chain(adc, S, C, A):
    Sp, Cp = next_pow2(S), next_pow2(C)
    wr, wd = hann(S).astype(np.float64), hann(C).astype(np.float64)
    z = pkg.synth.unpack_iiqq(adc.reshape(C, A, 2 * S))                                  # [c][a][s]
    rs = np.fft.fft(z * wr, n=Sp, axis=-1).transpose(1, 2, 0)                             # [a][r][c]
    dc = np.fft.fft(rs * wd, n=Cp, axis=-1)                                               # [a][r][d]
    P = (np.abs(dc) ** 2).sum(0)
    Wr, Wd = GR + TR, GD + TD
    noise = np.zeros_like(P)
    for r in range(Sp):
        rr = np.arange(max(0, r - Wr), min(Sp, r + Wr + 1))
        gr = np.arange(max(0, r - GR), min(Sp, r + GR + 1))
        for d in range(Cp):
            dd = np.arange(d - Wd, d + Wd + 1) % Cp
            gd = np.arange(d - GD, d + GD + 1) % Cp
            train = np.ones((rr.size, dd.size), bool)
            train[np.ix_(np.isin(rr, gr), np.isin(dd, gd))] = False                       # only ever ADD training cells
            noise[r, d] = P[np.ix_(rr, dd)][train].sum() / train.sum()
    thr = ALPHA * noise
    mask = P > thr
    near = np.abs(P - thr) <= 1e-5 * thr
    nth = 64 if A <= 64 else next_pow2(A)
    hits = np.argwhere(mask)                                                              # sorted by (r, d)
    angle, tie, peak = [], [], []
    for r, d in hits:
        Y = np.abs(np.fft.fft(dc[:, r, d], n=nth)) ** 2
        k = int(np.argmax(Y))
        angle.append(k if k < nth // 2 else k - nth)
        top2 = np.sort(Y)[-2:]
        tie.append(bool(top2[0] >= top2[1] * (1 - 1e-4)))
        nb = [(r + i, (d + j) % Cp) for i in (-1, 0, 1) for j in (-1, 0, 1) if (i or j) and 0 <= r + i < Sp]
        nb = [(a, b) for a, b in nb if mask[a, b]]
        peak.append(all(P[r, d] > P[a, b] or (P[r, d] == P[a, b] and (r, d) < (a, b)) for a, b in nb))
    return dict(P=P, noise_at_hits=noise[mask], mask=np.packbits(mask), near=np.packbits(near), hits=hits.astype(np.int32),
                angle_bin=np.array(angle, np.int32), angle_tie=np.array(tie, bool), peak=np.array(peak, bool))


out = {"cases": np.array(CASES, np.int32), "cfar": np.array([GR, GD, TR, TD, ALPHA])}
for (S, C, A, F, cfg) in CASES:
    adc = pkg.synth.cube_batch(F, S, C, A, cfg=cfg, n_targets=4)
    out[f"adc_sha256_{S}x{C}x{A}"] = np.frombuffer(hashlib.sha256(adc.tobytes()).digest(), np.uint8)
    for f in range(F):
        for k, v in chain(adc[f], S, C, A).items():
            out[f"{k}_{S}x{C}x{A}_f{f}"] = v
        print(S, C, A, "frame", f, "hits", len(out[f"hits_{S}x{C}x{A}_f{f}"]))
path = os.path.join(ROOT, "tests", "golden", "chain_numpy.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
