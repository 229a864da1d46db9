"""C-ABI surface: the library builds, loads and exports every symbol the headers declare; host-side
argument checking works without a GPU; and without a GPU nothing computes (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mmw_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.api.load()
    declared = _declared("mmw_radar.h") + _declared("mmw_legacy.h")
    assert len(declared) >= 24
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert sorted(declared) == sorted(pkg.api.C_ABI_SYMBOLS)
    # the reference's entry point, C++ linkage: same mangled name as the reference's own object file
    assert hasattr(lib, pkg.api.LEGACY_MANGLED)
    out = subprocess.run(["nm", "-D", "--defined-only", pkg.api.library_path()], capture_output=True, text=True).stdout
    assert " T " + pkg.api.LEGACY_MANGLED in out


def test_reference_caller_links_against_library(pkg):
    obj = pkg.build.REF_CALLER_OBJ
    if not os.path.exists(obj):
        pytest.skip("oracle/_ref/cudaBenchMarking.o not built here")
    undefined = subprocess.run(["nm", "-u", obj], capture_output=True, text=True).stdout
    assert pkg.api.LEGACY_MANGLED in undefined            # what the unmodified reference caller needs ...
    assert pkg.build.build_dropin() and os.path.exists(pkg.build.DROPIN_BIN)   # ... and our library provides


def test_headers_compile_as_c_and_cxx(tmp_path):
    c_src = tmp_path / "t.c"
    c_src.write_text('#include "mmw_radar.h"\n#include "mmw_legacy.h"\nint main(void){mmw_config c; mmw_default_config(&c,4,2,1,1); return sizeof(mmw_detection)==24?0:1;}\n')
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(c_src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cxx = tmp_path / "t.cpp"
    cxx.write_text('#include "mmw_legacy.h"\n#include "mmw_radar.h"\nstatic_assert(sizeof(Complex_t)==16,"");\nint main(){return 0;}\n')
    r = subprocess.run(["/usr/bin/g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(cxx)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_shipped_acceleration_header_is_interface_compatible(pkg, tmp_path):
    """include/acceleration.h (our own restatement of the reference's header: same guard, Complex_t, Timer, prototype) is
    enough to compile a caller written against the reference's header, and yields the reference's mangled symbol; where
    the reference tree is present, its unmodified cudaBenchMarking.cpp compiles against OUR header."""
    inc = os.path.join(ROOT, "include")
    t = tmp_path / "caller.cpp"
    t.write_text('#include "acceleration.h"\n#include "mmw_legacy.h"\n'
                 'static_assert(sizeof(Complex_t) == 16, "layout");\n'
                 'double run(short *in, Complex_t *base) { Timer t; t.reset(); double a = 0, b = 0, c = 0, d = 0;\n'
                 '  double r = cudaProcessing(in, base, 102400, &a, &b, &c, &d); return r + t.elapsed(); }\n')
    obj = tmp_path / "caller.o"
    r = subprocess.run(["/usr/bin/g++", "-std=c++11", "-Wall", "-Werror", "-c", "-I", inc, "-o", str(obj), str(t)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    undefined = subprocess.run(["nm", "-u", str(obj)], capture_output=True, text=True).stdout
    assert pkg.api.LEGACY_MANGLED in undefined
    ref = "/root/reference/cudaBenchMarking.cpp"
    if os.path.exists(ref):
        # our header first: its include guard makes the reference's own `#include "acceleration.h"` a no-op
        w = tmp_path / "ref_caller.cpp"
        w.write_text(f'#include "{inc}/acceleration.h"\n#include "{ref}"\n')
        r = subprocess.run(["/usr/bin/g++", "-m64", "-O1", "-w", "-c", "-o", str(tmp_path / "ref_caller.o"), str(w)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        undefined = subprocess.run(["nm", "-u", str(tmp_path / "ref_caller.o")], capture_output=True, text=True).stdout
        assert pkg.api.LEGACY_MANGLED in undefined


def test_struct_layouts_match_header(pkg):
    assert pkg.api.DET_DTYPE.itemsize == 24
    assert [pkg.api.DET_DTYPE.fields[n][1] for n in pkg.api.DET_DTYPE.names] == [0, 4, 6, 8, 12, 16, 18, 20]
    assert C.sizeof(pkg.api.Config) == 13 * 4
    cfg = pkg.api.Config()
    pkg.api.load().mmw_default_config(C.byref(cfg), 512, 256, 12, 8)
    assert (cfg.n_samples, cfg.n_chirps, cfg.n_antennas, cfg.max_frames) == (512, 256, 12, 8)
    assert (cfg.cfar_guard_r, cfg.cfar_guard_d, cfg.cfar_train_r, cfg.cfar_train_d) == (2, 2, 8, 4)
    assert cfg.cfar_alpha == 15.0 and cfg.max_det_per_frame == 1024 and cfg.keep_doppler_cube == 0
    assert cfg.lambda_over_d == 2.0 and cfg.device == -1


@pytest.mark.parametrize("kw,needle", [
    (dict(n_samples=510), "multiple of 4"),
    (dict(n_chirps=255), "multiple of 2"),
    (dict(n_antennas=0), "n_antennas"),
    (dict(n_antennas=300), "n_antennas"),
    (dict(n_samples=2048), "range FFT length"),
    (dict(n_samples=16), "range FFT length"),
    (dict(n_chirps=2048), "Doppler FFT length"),
    (dict(max_frames=0), "max_frames"),
    (dict(cfar_train=(100, 4)), "CFAR window"),
    (dict(max_det_per_frame=100000), "max_det_per_frame"),
])
def test_argument_validation_needs_no_gpu(pkg, kw, needle):
    args = dict(n_samples=256, n_chirps=128, n_antennas=4, max_frames=2)
    args.update(kw)
    with pytest.raises(pkg.RadarError) as ei:
        pkg.RadarContext(**args)
    assert ei.value.code == pkg.api.MMW_ERR_ARG and needle in str(ei.value)


def test_no_cpu_fallback_without_gpu(pkg):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.RadarError) as ei:
        pkg.RadarContext(256, 128, 4, 2)
    assert ei.value.code == pkg.api.MMW_ERR_CUDA
    d = pkg.api.load().mmw_legacy_process_frame(np.zeros(102400, np.int16).ctypes.data, np.zeros(25600).ctypes.data, 102400, None)
    assert d == float(pkg.api.MMW_ERR_CUDA)


def test_product_does_not_import_oracle():
    pkg_dir = os.path.join(ROOT, "cuda-based-mmwave-radar-object-detection-acceleration_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle." not in txt.replace("oracle/_ref", "").replace("oracle/Makefile", "") or f == "build.py", f
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "cufft" not in txt.lower() or "no cufft" in txt.lower(), f


def test_detect_path_constants_match_the_header():
    """MMW_DETECT_* of include/mmw_radar.h and the ctypes mirror's DETECT_* are the same numbers"""
    import re

    import __graft_entry__ as entry

    pkg = entry.load_package()
    hdr = open(os.path.join(ROOT, "include", "mmw_radar.h")).read()
    vals = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define MMW_DETECT_(\w+) (\d+)", hdr)}
    assert vals == {"AUTO": pkg.api.DETECT_AUTO, "PER_CELL": pkg.api.DETECT_PER_CELL, "REFFT": pkg.api.DETECT_REFFT}
