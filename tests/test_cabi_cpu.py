"""CPU-side checks: the C-ABI library loads and exports every symbol include/raft_corr_b200.h declares,
host-only entry points behave, and the Python mirror keeps the reference's error behaviour.  No compute."""
import ctypes
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from raft_optical_flow_b200 import _cabi, build as rcb_build  # noqa: E402


@pytest.fixture(scope="module")
def lib():
    rcb_build.build()
    return _cabi.lib()


def header_symbols():
    with open(os.path.join(ROOT, "include", "raft_corr_b200.h")) as f:
        return sorted(set(re.findall(r"RCB_API\s+[\w\s\*]+?\b(rcb_\w+)\s*\(", f.read())))


def test_library_exports_every_declared_symbol(lib):
    syms = header_symbols()
    assert len(syms) == 27
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_cabi.SIGNATURES) == syms  # the ctypes table binds exactly the header


def test_status_strings(lib):
    assert lib.rcb_abi_version() == 9
    assert lib.rcb_status_string(0) == b"ok"
    assert b"invalid" in lib.rcb_status_string(-1)
    assert b"unsupported" in lib.rcb_status_string(-2)


@pytest.mark.parametrize("H,W,want_h,want_w", [(55, 128, [55, 27, 13, 6], [128, 64, 32, 16]),
                                               (47, 156, [47, 23, 11, 5], [156, 78, 39, 19]),
                                               (46, 62, [46, 23, 11, 5], [62, 31, 15, 7])])
def test_pyramid_layout_floor_halving_and_padding(lib, H, W, want_h, want_w):
    lay = _cabi.pyramid_layout(3, H, W, 4)
    assert list(lay.H) == want_h and list(lay.W) == want_w  # core/corr.py:52-54 floor mode
    assert lay.tile_w == 4
    for l in range(4):
        assert lay.tiles_x[l] == -(-lay.W[l] // 4) and lay.tiles_y[l] == -(-lay.H[l] // 4)  # 4x4 tiles of 64 bytes
        assert lay.plane_stride[l] == lay.tiles_x[l] * lay.tiles_y[l] * 16
        assert lay.level_bytes[l] == 3 * H * W * lay.plane_stride[l] * 4
    half = _cabi.pyramid_layout(3, H, W, 4, _cabi.F16)
    assert half.tile_w == 8 and all(half.tiles_x[l] == -(-half.W[l] // 8) for l in range(4))


def test_layout_rejects_bad_arguments(lib):
    lay = _cabi.PyramidLayout()
    assert lib.rcb_pyramid_layout_query(1, 8, 8, 5, 0, ctypes.byref(lay)) == -2  # > RCB_MAX_LEVELS
    assert lib.rcb_pyramid_layout_query(0, 8, 8, 4, 0, ctypes.byref(lay)) == -1
    assert lib.rcb_pyramid_layout_query(1, 4, 4, 4, 0, ctypes.byref(lay)) == -1  # pooled away (4->2->1->0)


def test_entry_points_validate_before_touching_the_device(lib):
    null = _cabi.ptr_array([0, 0, 0, 0])
    assert lib.rcb_corr_lookup(null, None, None, 1, 8, 8, 4, 4, 0, None) == -1
    assert lib.rcb_corr_build(None, None, null, 1, 8, 8, 8, 4, 0, 0, None, 0, None) == -1
    assert lib.rcb_altcorr_forward(None, None, None, None, 1, 1, 8, 8, 8, 8, 8, 4, None) == -1
    assert lib.rcb_corr_lookup_plan_bytes() >= 16 * 128  # four tensor maps per level
    assert lib.rcb_corr_lookup_plan_init(None, 0, null, 1, 8, 8, 4, 4, 0) == -1
    assert lib.rcb_corr_lookup_planned(None, None, None, None) == -1
    blob = (ctypes.c_char * (lib.rcb_corr_lookup_plan_bytes() + 64))()
    plan = (ctypes.addressof(blob) + 63) & ~63
    assert lib.rcb_corr_lookup_planned(plan, ctypes.addressof(blob), ctypes.addressof(blob), None) == -1  # not initialised
    buf = (ctypes.c_float * 64)()
    p = ctypes.addressof(buf)
    assert lib.rcb_altcorr_forward(p, p, p, p, 1, 1, 2, 2, 2, 2, 4, 9, None) == -2  # radius > RCB_MAX_RADIUS


def test_python_mirror_has_no_cpu_path():
    from raft_optical_flow_b200 import AlternateCorrBlock, CorrBlock, alt_cuda_corr
    f = torch.zeros(1, 8, 8, 8)
    with pytest.raises(RuntimeError, match="fmap1 must be a CUDA tensor"):  # correlation.cpp:19
        CorrBlock(f, f)
    with pytest.raises(RuntimeError, match="fmap1 must be a CUDA tensor"):
        AlternateCorrBlock(f, f)
    with pytest.raises(RuntimeError, match="fmap1 must be a CUDA tensor"):
        alt_cuda_corr.forward(f.permute(0, 2, 3, 1).contiguous(), f, torch.zeros(1, 1, 8, 8, 2), 4)


def test_dropin_modules_resolve_like_the_reference_imports():
    """core/raft.py:8 does `from corr import CorrBlock, AlternateCorrBlock`; core/corr.py:6 `import alt_cuda_corr`."""
    import importlib
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    try:
        for m in ("corr", "alt_cuda_corr"):
            sys.modules.pop(m, None)
        corr = importlib.import_module("corr")
        ext = importlib.import_module("alt_cuda_corr")
        assert corr.CorrBlock.__init__.__code__.co_varnames[:5] == ("self", "fmap1", "fmap2", "num_levels", "radius")
        assert callable(corr.CorrBlock.corr) and callable(ext.forward) and callable(ext.backward)
    finally:
        sys.path.remove(os.path.join(ROOT, "dropin"))
        for m in ("corr", "alt_cuda_corr"):
            sys.modules.pop(m, None)


def test_convc1_pack_size_and_argument_checks(lib):
    """Host logic of the fused lookup + convc1 entry points (include/raft_corr_b200.h): K is 96 entries per level at
    r = 4 and 64 at r = 3 (window rows padded to even length, levels to a multiple of 16), rows padded to 128."""
    assert lib.rcb_corr_convc1_pack_bytes(256, 4, 4) == 256 * 4 * 96 * 2   # BasicMotionEncoder.convc1 (core/update.py:182)
    assert lib.rcb_corr_convc1_pack_bytes(96, 4, 3) == 128 * 4 * 64 * 2    # SmallMotionEncoder.convc1 (core/update.py:136)
    assert lib.rcb_corr_convc1_pack_bytes(16, 3, 4) == 128 * 5 * 64 * 2      # 144 columns -> five 32-column chunks
    for cout, levels, radius in ((24, 4, 4), (272, 4, 4), (0, 4, 4), (256, 1, 4), (256, 5, 4), (256, 4, 2), (256, 4, 5)):
        assert lib.rcb_corr_convc1_pack_bytes(cout, levels, radius) == 0
    buf = (ctypes.c_float * 64)()
    p = ctypes.addressof(buf)
    assert lib.rcb_corr_convc1_pack(None, p, 256, 4, 4, None) == -1
    assert lib.rcb_corr_convc1_pack(p, p, 24, 4, 4, None) == -2
    assert lib.rcb_corr_lookup_convc1(None, p, p, None, p, 256, 1, None) == -1
    blob = (ctypes.c_char * (lib.rcb_corr_lookup_plan_bytes() + 64))()
    plan = (ctypes.addressof(blob) + 63) & ~63
    assert lib.rcb_corr_lookup_convc1(plan, p, p, None, p, 256, 1, None) == -1  # plan not initialised
