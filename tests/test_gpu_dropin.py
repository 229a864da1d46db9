"""Drop-in proof: the reference's UNMODIFIED cudaBenchMarking.cpp (compiled from /root/reference into
oracle/_ref/cudaBenchMarking.o, linked against libmmw_radar_b200.so) runs its cpuTiming()/cudaTiming()
loops on a synthetic fhy_direct.bin."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu


def test_unmodified_reference_caller_runs_against_our_library(pkg, tmp_path):
    exe = pkg.build.DROPIN_BIN
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin_acceleration was not built (needs /root/reference at build time)")
    cap = pkg.synth.legacy_capture(90, seed=0)
    cap.tofile(tmp_path / "fhy_direct.bin")
    assert os.path.getsize(tmp_path / "fhy_direct.bin") == 18_432_000
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    assert out.count("Inner CUDA Timing") == 89                     # one line per frame, as acceleration.cu:533
    assert "Total Time for 89 frames" in out and "cuda totalTime" in out and "cuda inner time" in out
    print("\n" + "\n".join(l for l in out.splitlines() if "Inner CUDA" not in l))


def test_verification_harness_of_the_reference_main(pkg, orc, tmp_path):
    """the CPU-vs-GPU check the reference left commented out (cudaBenchMarking.cpp:405-419,
    tolerance 1e-5), as a C++ program calling cudaProcessing() the way cudaTiming() does."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "verify_legacy"
    cmd = ["/usr/bin/g++", "-O2", "-o", str(exe), os.path.join(root, "tests", "verify_legacy.cpp"),
           "-I", os.path.join(root, "include"), "-I", os.path.join(root, "oracle"),
           pkg.api.library_path(), orc.ORACLE_SO, "-Wl,-rpath," + os.path.dirname(pkg.api.library_path()),
           "-Wl,-rpath," + os.path.dirname(orc.ORACLE_SO)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for seed, moving in ((0, False), (7, True)):
        cap = pkg.synth.legacy_capture(40, seed=seed, moving=moving)
        path = tmp_path / f"cap{seed}.bin"
        cap.tofile(path)
        r = subprocess.run([str(exe), str(path)], capture_output=True, text=True, timeout=300,
                           env=dict(os.environ, MMW_LEGACY_QUIET="1"))
        assert r.returncode == 0, r.stdout + r.stderr
        assert "verified 39 frames, 0 mismatches" in r.stdout


def test_cxx_host_program_against_the_c_abi(pkg, tmp_path):
    """examples/b200_timing.cpp: plain C++ over include/mmw_radar.h (capture-file ingest + physical units), the program
    INTEGRATION.md describes; its detection count must equal the in-memory path's."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "b200_timing"
    r = subprocess.run(["/usr/bin/g++", "-O2", "-std=c++11", "-Wall", "-Werror", "-o", str(exe), os.path.join(root, "examples", "b200_timing.cpp"),
                        "-I", os.path.join(root, "include"), pkg.api.library_path(),
                        "-Wl,-rpath," + os.path.dirname(pkg.api.library_path())], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    S, C, A = 100, 128, 4
    adc = pkg.synth.cube_batch(9, S, C, A, cfg=12, n_targets=3)
    adc.tofile(tmp_path / "cap.bin")
    with pkg.RadarContext(S, C, A, 9) as ctx:
        want, _ = ctx.process_host(adc, 9)
    r = subprocess.run([str(exe), str(tmp_path / "cap.bin"), str(S), str(C), str(A), "4"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"b200 detections {len(want)}," in r.stdout and "(9 frames" in r.stdout and "range " in r.stdout
    r = subprocess.run([str(exe), str(tmp_path / "missing.bin"), str(S), str(C), str(A)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 1 and "unable to read the specified file" in r.stdout


def test_cxx_streaming_program(pkg, tmp_path):
    """examples/b200_stream.cpp: the per-frame loop over mmw_submit_host / mmw_wait with a ring of contexts; every depth
    gives the detection count of the batched in-memory path"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "b200_stream"
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    r = subprocess.run(["/usr/bin/g++", "-O2", "-std=c++11", "-Wall", "-Werror", "-o", str(exe), os.path.join(root, "examples", "b200_stream.cpp"),
                        "-I", os.path.join(root, "include"), "-I", os.path.join(cuda, "include"), pkg.api.library_path(),
                        "-L", os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath," + os.path.dirname(pkg.api.library_path())],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    S, C, A = 256, 128, 12
    adc = pkg.synth.cube_batch(11, S, C, A, cfg=5, n_targets=4)
    adc.tofile(tmp_path / "cap.bin")
    with pkg.RadarContext(S, C, A, 11) as ctx:
        want, _ = ctx.process_host(adc, 11)
    for depth in (1, 2, 4):
        r = subprocess.run([str(exe), str(tmp_path / "cap.bin"), str(S), str(C), str(A), str(depth)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        assert f"b200 stream detections {len(want)} in 11 frames" in r.stdout, r.stdout
