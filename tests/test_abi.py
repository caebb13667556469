"""The C-ABI library loads without a GPU and exports every symbol include/gsr.h declares; the C++ shim exports
the reference's exact mangled CudaRasterizer::Rasterizer symbols. No compute call is made here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gsr.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"GSR_API\s+[\w\s\*]+?\b(gsr_\w+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for required in ("gsr_forward", "gsr_backward", "gsr_mark_visible", "gsr_required_geom", "gsr_required_image",
                     "gsr_required_binning", "gsr_last_error", "gsr_view_geom", "gsr_view_binning", "gsr_view_image"):
        assert required in syms


def test_library_exports_every_declared_symbol(built_library):
    from gs_livm_b200 import _lib
    for name in declared_symbols():
        assert hasattr(built_library, name), f"{name} declared in include/gsr.h but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes prototype in _lib.SYMBOLS"
    from gs_livm_b200 import _lib
    assert built_library.gsr_abi_version() == _lib.ABI_VERSION == 5


def test_ctypes_struct_layout_matches_the_library(built_library):
    from gs_livm_b200 import _lib
    for which, st in enumerate((_lib.FwdArgs, _lib.BwdArgs, _lib.ArrayView, _lib.GeomView, _lib.BinningView,
                                _lib.ImageView, _lib.AdamGroup, _lib.WindowComm, _lib.VgprParams, _lib.AsyncStatus,
                                _lib.WindowRows)):
        assert built_library.gsr_sizeof(which) == C.sizeof(st), st.__name__
    assert built_library.gsr_sizeof(99) == 0


def test_cpp_shim_exports_reference_mangled_names(built_library):
    """Exactly the symbols the reference's src/gs/rasterize_points.cu leaves undefined (nm -u)."""
    from gs_livm_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mangled in (
        "_ZN14CudaRasterizer10Rasterizer11markVisibleEiPfS1_S1_Pb",
        "_ZN14CudaRasterizer10Rasterizer7forwardESt8functionIFPcmEES4_S4_iiiPKfiiS6_S6_S6_S6_S6_fS6_S6_S6_S6_S6_ffbPfS7_S7_Pib",
        "_ZN14CudaRasterizer10Rasterizer8backwardEiiiiPKfiiS2_S2_S2_S2_fS2_S2_S2_S2_S2_ffPKiPcS5_S5_S2_S2_PfS6_S6_S6_S6_S6_S6_S6_S6_b",
    ):
        assert mangled in out


def test_required_sizes_are_monotonic_and_cover_the_layout(built_library):
    lib = built_library
    g1, g2 = lib.gsr_required_geom(1000), lib.gsr_required_geom(2000)
    assert 0 < g1 < g2 and g2 >= 2000 * (48 + 4 + 4 + 4 + 1)
    assert lib.gsr_required_image(752, 480) >= 752 * 480 * 8 + 1410 * 16
    b1, b2 = lib.gsr_required_binning(1000), lib.gsr_required_binning(100000)
    assert 0 < b1 < b2 and b2 >= 100000 * (48 + 8 + 8 + 4 + 4)
    assert lib.gsr_required_geom(-1) == 0 and lib.gsr_required_binning(-5) == 0


def test_argument_errors_are_reported_not_thrown(built_library):
    """Return codes + gsr_last_error(), never an exception across the C ABI; validation happens before any CUDA
    call, so this runs without a GPU."""
    from gs_livm_b200 import _lib
    lib = built_library
    n = C.c_int(7)
    cb = _lib.ALLOC_FN(lambda ctx, size: 0)
    assert lib.gsr_forward(None, cb, None, cb, None, cb, None, None, C.byref(n)) == -1
    assert b"null" in lib.gsr_last_error()
    a = _lib.FwdArgs()
    a.P, a.width, a.height = 0, 16, 16
    assert lib.gsr_forward(C.byref(a), cb, None, cb, None, cb, None, None, C.byref(n)) == -1
    assert n.value == 0
    a.P = 10  # still missing every pointer
    assert lib.gsr_forward(C.byref(a), cb, None, cb, None, cb, None, None, C.byref(n)) == -1
    assert lib.gsr_backward(None, None) == -1
    assert lib.gsr_mark_visible(-1, None, None, None, None, None) == -1
    assert lib.gsr_mark_visible(0, None, None, None, None, None) == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from gs_livm_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load()


def test_configure_knobs_round_trip_without_a_gpu(built_library):
    """gsr_configure / gsr_config_get (host-only state): values are clamped, split_len is rounded to 2^k - 1, a
    negative value restores the default, an unknown knob is an error."""
    lib = built_library
    got = C.c_int(0)
    assert lib.gsr_config_get(b"split_len", C.byref(got)) == 0 and got.value == 2047
    assert lib.gsr_configure(b"split_len", 1500) == 0
    assert lib.gsr_config_get(b"split_len", C.byref(got)) == 0 and got.value == 1023
    assert lib.gsr_configure(b"split_len", 1) == 0
    assert lib.gsr_config_get(b"split_len", C.byref(got)) == 0 and got.value == 255
    assert lib.gsr_configure(b"split_len", -1) == 0
    assert lib.gsr_config_get(b"split_len", C.byref(got)) == 0 and got.value == 2047
    assert lib.gsr_configure(b"bucket_target", 10) == 0
    assert lib.gsr_config_get(b"bucket_target", C.byref(got)) == 0 and got.value == 256
    assert lib.gsr_configure(b"bucket_target", -1) == 0
    assert lib.gsr_configure(b"no_such_knob", 1) != 0 and b"unknown knob" in lib.gsr_last_error()
