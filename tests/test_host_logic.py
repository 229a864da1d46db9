"""Host-side logic that needs no GPU: detections -> physical units (mmw_to_physical), the oracle's base-frame
subtraction, and argument checking of the capture-file / legacy-file entry points."""
import ctypes as C
import os

import numpy as np
import pytest


def test_default_radar_params_are_the_reference_constants(pkg):
    rp = pkg.api.default_radar_params()            # cudaBenchMarking.cpp:10-19
    assert (rp.f0_hz, rp.slope_hz_per_s, rp.fs_hz, rp.chirp_period_s, rp.light_speed) == (77e9, 5.987e12, 2.0e6, 64e-6, 3.0e8)


def test_to_physical_matches_closed_forms(pkg):
    Sp, Cp = 512, 256
    d = np.zeros(4, pkg.DET_DTYPE)
    d["frame"] = [0, 1, 2, 3]
    d["range_bin"] = [0, 10, 255, 511]
    d["doppler_bin"] = [0, 1, 128, 255]              # 128 wraps to -128, 255 to -1
    d["power"] = [100.0, 1000.0, 50.0, 0.0]
    d["noise"] = [1.0, 10.0, 50.0, 1.0]
    d["angle_rad"] = [0.0, np.pi / 6, -np.pi / 2, 0.25]
    d["flags"] = [1, 0, 1, 0]
    t = pkg.api.to_physical(d, Sp, Cp)
    c, f0, mu, fs, tr = 3.0e8, 77e9, 5.987e12, 2.0e6, 64e-6
    want_r = c * (d["range_bin"].astype(np.float64) / Sp * fs) / (2 * mu)
    dw = np.array([0, 1, -128, -1], np.float64)
    want_v = 0.5 * (c / f0) * dw / (Cp * tr)
    assert np.allclose(t["range_m"], want_r, rtol=1e-6) and np.allclose(t["velocity_mps"], want_v, rtol=1e-6)
    assert np.allclose(t["angle_deg"], [0, 30, -90, np.degrees(0.25)], atol=1e-4)
    assert np.allclose(t["snr_db"], [20, 20, 0, 0], atol=1e-5)
    assert list(t["frame"]) == [0, 1, 2, 3] and list(t["flags"]) == [1, 0, 1, 0]


def test_to_physical_reproduces_the_reference_distance_formula(pkg):
    """One chirp of the reference config: bin k of a 128-point (padded from 100) range FFT at Fs = 2 MHz is the same
    beat frequency the reference's flat-FFT formula (cudaBenchMarking.cpp:301-303) assigns to raw bin 128 k."""
    d = np.zeros(1, pkg.DET_DTYPE)
    d["range_bin"] = 15
    d["power"], d["noise"] = 2.0, 1.0
    got = float(pkg.api.to_physical(d, 128, 128)["range_m"][0])
    fs, c, mu, n_ext, n_valid = 2.0e6, 3.0e8, 5.987e12, 16384, 12800
    raw = 15 * 128                                   # same beat frequency in the 16 384-point spectrum
    fs_ext = fs * n_ext / n_valid
    ref = c * ((raw * n_valid // n_ext) / n_ext * fs_ext) / (2 * mu)
    assert abs(got - ref) / ref < 1e-6


def test_to_physical_rejects_bad_arguments(pkg):
    L = pkg.api.load()
    rp = pkg.api.default_radar_params()
    out = np.empty(1, pkg.api.TARGET_DTYPE)
    d = np.zeros(1, pkg.DET_DTYPE)
    assert L.mmw_to_physical(None, 64, 64, d.ctypes.data_as(C.c_void_p), 1, out.ctypes.data_as(C.c_void_p)) == pkg.api.MMW_ERR_ARG
    rp.fs_hz = 0.0
    assert L.mmw_to_physical(C.byref(rp), 64, 64, d.ctypes.data_as(C.c_void_p), 1, out.ctypes.data_as(C.c_void_p)) == pkg.api.MMW_ERR_ARG
    assert b"positive" in L.mmw_last_error()
    assert pkg.api.to_physical(np.zeros(0, pkg.DET_DTYPE), 64, 64).size == 0


def test_oracle_base_frame_subtraction(orc, pkg):
    """process_frames(base=b) == process_frames(adc - b) wherever the difference fits int16."""
    S, C_, A, F = 64, 64, 2, 3
    adc = pkg.synth.cube_batch(F, S, C_, A, cfg=5, n_targets=3, noise_sigma=20.0) // 2
    base = pkg.synth.cube(99, S, C_, A, cfg=5, n_targets=1, noise_sigma=20.0) // 2
    wr, wd = orc.hann_periodic(S), orc.hann_periodic(C_)
    a = orc.process_frames(adc, F, S, C_, A, wr, wd, want=("rs", "P"), base=base)
    b = orc.process_frames((adc - base[None, :]).astype(np.int16), F, S, C_, A, wr, wd, want=("rs", "P"))
    assert np.array_equal(a["rs"], b["rs"]) and np.array_equal(a["P"], b["P"])
    assert a["dets"].tobytes() == b["dets"].tobytes()
    # a frame identical to the base frame vanishes
    z = orc.process_frames(base[None, :], 1, S, C_, A, wr, wd, want=("P",), base=base)
    assert not z["P"].any() and len(z["dets"]) == 0


def test_file_entry_points_report_a_missing_file(pkg, tmp_path):
    L = pkg.api.load()
    n = C.c_int(-1)
    dist = np.empty(4, np.float64)
    rc = L.mmw_legacy_process_file(os.fsencode(str(tmp_path / "nope.bin")), dist.ctypes.data_as(C.c_void_p), None, 4, C.byref(n))
    assert rc == pkg.api.MMW_ERR_ARG and n.value == 0
    assert b"unable to read the specified file" in L.mmw_last_error()      # the reference's message, cudaBenchMarking.cpp:346
    rc = L.mmw_process_capture_file(None, b"x", 0, 0, 0, None, 0, None, None)
    assert rc == pkg.api.MMW_ERR_ARG


def test_legacy_distance_from_raw_is_the_reference_formula(pkg, orc):
    for raw in (0, 1, 127, 1966, 6552):
        assert pkg.api.legacy_distance_from_raw(raw) == orc.lib().orc_distance_from_raw(raw, 12800, 16384)


def test_register_butterfly_networks_against_a_direct_dft(tmp_path):
    """csrc/fft_regs.cuh compiled for the HOST (its packed arithmetic spelled with fmaf): every radix the kernels use,
    impulse / the reference's ramp vector (acceleration.cu:361-365) / random int16-range inputs, against a direct fp64 DFT
    with the reference's sign convention (cudaBenchMarking.cpp:88-104).  No GPU involved: nvcc is only the compiler."""
    import shutil
    import subprocess

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "dft_regs_host_check")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([nvcc, "-ccbin", cxx, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-o", exe,
                    os.path.join(root, "tests", "dft_regs_host_check.cu")], check=True, capture_output=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    errs = {int(l.split()[0]): float(l.split()[1]) for l in res.stdout.splitlines()}
    assert res.returncode == 0 and sorted(errs) == [2, 4, 8, 16, 32], res.stdout + res.stderr
    assert max(errs.values()) < 2e-6      # tolerance: a few fp32 ulps of the largest output bin (measured 1.1e-7 at radix 32)
