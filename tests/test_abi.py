"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/msdfgen_b200.h declares,
and refuses to compute without a GPU (no silent fallback)."""
import ctypes as C
import os
import re

import pytest

from msdfgen_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "msdfgen_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(b200msdf_[a-z_0-9]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    lib = _lib.load()
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert sorted(_lib.EXPORTS) == declared


def test_host_library_exports_every_declared_symbol():
    """include/msdfgen_b200_host.h (batch preparation, font ingest) against libmsdfgen_b200_host.so, and the header compiles as C."""
    import subprocess
    from msdfgen_b200 import prep
    path = os.path.join(ROOT, "include", "msdfgen_b200_host.h")
    declared = sorted(set(re.findall(r"\b(b200host_[a-z_0-9]+)\s*\(", open(path).read())))
    assert len(declared) >= 14
    lib = prep.host_lib()
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    exported = subprocess.run(["nm", "-D", "--defined-only", prep.HOST_LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert sorted(set(re.findall(r"\b(b200host_[a-z_0-9]+)\b", exported))) == declared
    subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", path], check=True, capture_output=True)


def test_version_and_default_config():
    lib = _lib.load()
    assert lib.b200msdf_version() == 101
    cfg = _lib.Config()
    lib.b200msdf_default_config(C.byref(cfg))
    # MSDFGeneratorConfig() defaults, /root/reference/core/generator-config.h:46,52,62
    assert (cfg.overlap_support, cfg.ec_mode, cfg.ec_distance_check_mode) == (1, 2, 1)
    assert cfg.min_deviation_ratio == cfg.min_improve_ratio == 1.11111111111111111
    # the scanline pass is off by default; distanceSignCorrection's defaults (core/rasterization.h:14)
    assert (cfg.scanline_pass, cfg.fill_rule, cfg.sdf_zero_value) == (0, 0, 0.5)


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.Transform) == 48
    assert C.sizeof(_lib.Bitmap) == 24
    assert C.sizeof(_lib.Shapes) == 64
    assert C.sizeof(_lib.Config) == 56


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.b200msdf_create(0, C.byref(h))
    assert rc == -2 and not h.value
    assert b"no CPU fallback" in lib.b200msdf_last_error()
    # compute entry points reject a null context instead of computing anything
    assert lib.b200msdf_generate(None, 3, None, None, None, None) != 0


def test_product_does_not_reference_the_oracle():
    """The shipped library and package must not import, link or load anything under oracle/ (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "msdfgen_b200")
    banned = ("libmsdf_oracle", "libmsdfgen_ref", "from oracle", "import oracle", "oracle.bindings", "msdf_oracle.c", "ref_wrapper")
    checked = 0
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                checked += 1
                for b in banned:
                    assert b not in text, (os.path.join(dirpath, f), b)
    assert checked > 10
    # and the built libraries do not depend on the oracle's shared objects
    import subprocess
    for lib in ("libmsdfgen_b200.so", "libmsdfgen_b200_host.so"):
        out = subprocess.run(["ldd", os.path.join(pkg, lib)], capture_output=True, text=True).stdout
        assert "oracle" not in out and "msdfgen_ref" not in out
