"""Drop-in boundary test: one msdfgen USER program (tests/dropin/caller.cpp, public msdfgen.h API only) is compiled against
the unmodified reference and against the B200 host mirror; the CPU leg checks that both compile and that the reference
build agrees with the oracle port, the GPU leg checks the B200 build against it."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from util import compare, to_bytes
from msdfgen_b200.synth import synth_glyph

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "dropin", "caller.cpp")
OUT = os.path.join(ROOT, "tests", "dropin", "_build")
REF = "/root/reference"
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _build_ref():
    if not os.path.isdir(REF):
        pytest.skip("/root/reference is not present on this box")
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libcaller_ref.so")
    if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(SRC):
        core = sorted(os.path.join(REF, "core", f) for f in os.listdir(os.path.join(REF, "core")) if f.endswith(".cpp"))
        subprocess.run([CXX, "-O2", "-ffp-contract=off", "-std=c++11", "-pthread", "-fPIC", "-shared", "-DMSDFGEN_PUBLIC=", "-DMSDFGEN_USE_CPP11",
                        f"-I{REF}", "-o", lib, SRC] + core, check=True, capture_output=True)
    return C.CDLL(lib)


def _build_b200():
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libcaller_b200.so")
    host = os.path.join(ROOT, "msdfgen_b200")
    deps = [SRC, os.path.join(host, "host", "msdfgen.h"), os.path.join(host, "libmsdfgen_b200_host.so")]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(d) for d in deps):
        subprocess.run([CXX, "-O2", "-ffp-contract=off", "-std=c++11", "-pthread", "-fPIC", "-shared", f"-I{host}/host", "-o", lib, SRC,
                        f"-L{host}", "-lmsdfgen_b200_host", "-lmsdfgen_b200", f"-Wl,-rpath,{host}"], check=True, capture_output=True)
    return C.CDLL(lib)


def _args(g):
    return (C.c_int(g.n_contours), g.cbegin.ctypes.data_as(C.POINTER(C.c_uint32)), g.points.ctypes.data_as(C.POINTER(C.c_double)),
            g.orders.ctypes.data_as(C.POINTER(C.c_uint8)))


def _run(lib, g, w=24, h=20):
    res = {}
    a = np.full((h, w, 3), np.nan, np.float32)
    res["edges"] = lib.caller_msdf(*_args(g), C.c_int(0), C.c_int(w), C.c_int(h), a.ctypes.data_as(C.POINTER(C.c_float)))
    res["msdf"] = a
    b = np.full((4, h, w, 4), 7.0, np.float32)
    assert lib.caller_all_modes(*_args(g), C.c_int(w), C.c_int(h), C.c_int(1), b.ctypes.data_as(C.POINTER(C.c_float))) == 0
    res["modes"] = b
    c = np.full((3, h, w, 3), np.nan, np.float32)
    st = np.zeros((h, w), np.uint8)
    assert lib.caller_error_correction(*_args(g), C.c_int(w), C.c_int(h), c.ctypes.data_as(C.POINTER(C.c_float)), st.ctypes.data_as(C.POINTER(C.c_uint8))) == 0
    res["ec"], res["stencil"] = c, st
    d = np.full(h * w * 9, np.nan, np.float32)
    assert lib.caller_scanline(*_args(g), C.c_int(w), C.c_int(h), d.ctypes.data_as(C.POINTER(C.c_float))) == 0
    res["scanline"] = d
    e = np.full(4 * h * w * 3 + h * w, np.nan, np.float32)
    assert lib.caller_render(*_args(g), C.c_int(w), C.c_int(h), e.ctypes.data_as(C.POINTER(C.c_float))) == 0
    res["render"] = e
    f = np.full(3, np.nan, np.float64)
    assert lib.caller_estimate(*_args(g), C.c_int(w), C.c_int(h), f.ctypes.data_as(C.POINTER(C.c_double))) == 0
    res["sdferr"] = f
    return res


def test_caller_compiles_against_both_interfaces():
    """Compile-time drop-in: the same user source builds against the reference headers and against the B200 mirror."""
    _build_b200()
    _build_ref()


def _model(lib, g, raw):
    cap = 400 * (len(g.orders) + 8)
    out = np.full(cap, np.nan, np.float64)
    n = lib.caller_shape_model(*_args(g), C.c_int(raw), out.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(cap))
    assert n > 0
    return out[:n].view(np.uint64)


def test_shape_object_model_matches_reference(ref):
    """The host mirror's Shape / Contour / EdgeSegment objects answer every geometric query of the msdfgen API (point, direction,
    directionChange, bound, length, splitInThirds, convertToCubic, move*Point, reverse, winding, boundMiters, getBounds, normalize,
    both colourings) with the reference's doubles, bit for bit — on ordinary glyphs, on cusps / teardrops / degenerate curves, and
    with edges built past the factory (raw=1: no order lowering).  No device involved."""
    from test_host_model import _cusp_shapes
    from util import fuzz_shape
    from msdfgen_b200.flat import FlatShapes
    mine, theirs = _build_b200(), _build_ref()
    shapes = [synth_glyph(200 + i, cubic_fraction=0.4 if i % 2 else 0.0) for i in range(12)] + _cusp_shapes()
    shapes += [FlatShapes.single(*fuzz_shape(seed)[1:5]) for seed in range(0, 108, 5)]
    for g in shapes:
        for raw in (0, 1):
            a, b = _model(mine, g, raw), _model(theirs, g, raw)
            assert a.shape == b.shape and np.array_equal(a, b), (raw, int(np.argmax(a != b)) if a.shape == b.shape else (a.shape, b.shape))


def test_reference_build_runs(ref):
    lib = _build_ref()
    r = _run(lib, synth_glyph(123, cubic_fraction=0.2))
    assert r["edges"] > 0 and np.isfinite(r["msdf"]).all() and (r["stencil"] & 2).all()


@pytest.mark.gpu
def test_b200_build_matches_reference_build():
    """Run-time drop-in on the GPU: same caller, two backends, same texels (falls back to a fixture when the reference
    tree is not on the box)."""
    b200 = _build_b200()
    fixture = os.path.join(ROOT, "tests", "golden", "dropin_v1.npz")
    z = np.load(fixture)
    for i, seed in enumerate(z["seeds"]):
        g = synth_glyph(int(seed), cubic_fraction=0.2)
        mine = _run(b200, g)
        assert mine["edges"] == int(z[f"edges{i}"])
        for key in ("msdf", "modes", "ec", "scanline", "render"):
            want = z[f"{key}{i}"]
            bits, max_abs, over = compare(mine[key], want)
            assert over == 0 and max_abs <= 1e-4, (key, seed, bits, max_abs)
            assert np.array_equal(to_bytes(mine[key]), to_bytes(want))
        assert np.array_equal(mine["stencil"], z[f"stencil{i}"])
        assert np.array_equal(mine["sdferr"], z[f"sdferr{i}"]), (mine["sdferr"], z[f"sdferr{i}"])


def _threads(lib, g, n_threads, reps, w=24, h=20):
    sizes = [(w + t, h + t) for t in range(n_threads)]
    out = np.full(sum(a * b * 3 for a, b in sizes), np.nan, np.float32)
    calls = lib.caller_threads(C.c_int(n_threads), C.c_int(reps), *_args(g), C.c_int(w), C.c_int(h), out.ctypes.data_as(C.POINTER(C.c_float)))
    assert calls == n_threads * reps
    imgs, off = [], 0
    for a, b in sizes:
        imgs.append(out[off:off + a * b * 3].reshape(b, a, 3))
        off += a * b * 3
    return imgs


def test_reference_build_is_reentrant(ref):
    """The contract the mirror has to keep: concurrent generator calls from several host threads give what sequential calls give."""
    lib = _build_ref()
    g = synth_glyph(321, cubic_fraction=0.2)
    seq = [_threads(lib, g, t + 1, 1)[t] for t in range(4)]
    par = _threads(lib, g, 4, 3)
    for a, b in zip(seq, par):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.gpu
def test_b200_build_is_reentrant_and_matches_reference():
    """Four host threads call msdfgen::generateMSDF concurrently through the C++ mirror (one context per calling thread, each
    with its own stream): the images equal the single-threaded ones and the reference's."""
    b200 = _build_b200()
    g = synth_glyph(321, cubic_fraction=0.2)
    seq = [_threads(b200, g, t + 1, 1)[t] for t in range(4)]
    for rep in range(3):
        par = _threads(b200, g, 4, 8)
        for a, b in zip(seq, par):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    z = np.load(os.path.join(ROOT, "tests", "golden", "dropin_v1.npz"))
    for t in range(4):
        want = z[f"threads{t}"]
        bits, max_abs, over = compare(seq[t], want)
        assert over == 0 and max_abs <= 1e-4 and np.array_equal(to_bytes(seq[t]), to_bytes(want)), (t, bits, max_abs)


@pytest.mark.gpu
def test_two_threads_share_one_context_serialised(ctx):
    """A context is not internally locked (include/msdfgen_b200.h: one per thread); two threads may still share one if the caller
    serialises the calls — e.g. a worker pool around one GPU.  ctypes releases the GIL during the calls, so the lock is real."""
    import threading
    from msdfgen_b200 import api, prep
    from msdfgen_b200.synth import synth_atlas, glyph_transform
    atlases = [prep.prepare(synth_atlas(64, seed=500 + i)) for i in range(2)]
    xf = glyph_transform(32)
    want = []
    for a in atlases:
        o = np.empty((64, 32, 32, 3), np.float32)
        ctx.generate_batch(api.MSDF, o, a, xf)
        want.append(o)
    lock = threading.Lock()
    got = [np.empty((64, 32, 32, 3), np.float32) for _ in range(2)]
    errors = []

    def worker(i):
        try:
            for _ in range(10):
                with lock:
                    ctx.generate_batch(api.MSDF, got[i], atlases[i], xf)
                if not np.array_equal(got[i].view(np.uint32), want[i].view(np.uint32)):
                    errors.append(i)
        except Exception as e:          # noqa: BLE001
            errors.append(repr(e))
    ts = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
