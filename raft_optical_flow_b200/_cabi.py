"""ctypes binding of libraftcorr_b200.so (include/raft_corr_b200.h).

The library is the product: if it is missing or fails to load, importing the operators raises --
there is no CPU or PyTorch fallback.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libraftcorr_b200.so")
if os.environ.get("RCB_USE_DEBUG_LIB") == "1":  # timing / profiling hooks (tools/time_*.py); see build.py --debug
    LIB_PATH = os.path.join(HERE, "libraftcorr_b200_debug.so")
if os.environ.get("RCB_LIB_VARIANT"):  # A/B builds of build.py --variant (debug-hook builds with compile-time switches)
    LIB_PATH = os.path.join(HERE, f"libraftcorr_b200_{os.environ['RCB_LIB_VARIANT']}.so")

ABI_VERSION = 9
MAX_LEVELS = 4
MAX_RADIUS = 4

# enum rcb_dtype / rcb_build_mode
F32, F16 = 0, 1
BUILD_FP32_SIMT, BUILD_BF16X3, BUILD_BF16, BUILD_F16F8 = 0, 1, 2, 3
BUILD_MODES = {"fp32": BUILD_FP32_SIMT, "bf16x3": BUILD_BF16X3, "bf16": BUILD_BF16, "f16f8": BUILD_F16F8}

_vp = ctypes.c_void_p
_i = ctypes.c_int


class PyramidLayout(ctypes.Structure):
    _fields_ = [
        ("levels", ctypes.c_int32),
        ("dtype", ctypes.c_int32),
        ("tile_w", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("H", ctypes.c_int32 * MAX_LEVELS),
        ("W", ctypes.c_int32 * MAX_LEVELS),
        ("tiles_x", ctypes.c_int32 * MAX_LEVELS),
        ("tiles_y", ctypes.c_int32 * MAX_LEVELS),
        ("plane_stride", ctypes.c_int64 * MAX_LEVELS),
        ("level_bytes", ctypes.c_int64 * MAX_LEVELS),
    ]


# name -> (restype, argtypes); every symbol include/raft_corr_b200.h declares
SIGNATURES = {
    "rcb_abi_version": (_i, []),
    "rcb_status_string": (ctypes.c_char_p, [_i]),
    "rcb_pyramid_layout_query": (_i, [_i, _i, _i, _i, _i, ctypes.POINTER(PyramidLayout)]),
    "rcb_corr_build_workspace_bytes": (ctypes.c_size_t, [_i, _i, _i, _i, _i]),
    "rcb_corr_build": (_i, [_vp, _vp, ctypes.POINTER(_vp), _i, _i, _i, _i, _i, _i, _i, _vp, ctypes.c_size_t, _vp]),
    "rcb_corr_pack_fmaps": (_i, [_vp, _vp, ctypes.c_size_t, _i, _i, _i, _i, _i, _vp]),
    "rcb_corr_build_packed": (_i, [_vp, ctypes.c_size_t, ctypes.POINTER(_vp), _i, _i, _i, _i, _i, _i, _i, _vp]),
    "rcb_corr_lookup": (_i, [ctypes.POINTER(_vp), _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "rcb_corr_lookup_plan_bytes": (ctypes.c_size_t, []),
    "rcb_corr_lookup_plan_init": (_i, [_vp, ctypes.c_size_t, ctypes.POINTER(_vp), _i, _i, _i, _i, _i, _i]),
    "rcb_corr_lookup_planned": (_i, [_vp, _vp, _vp, _vp]),
    "rcb_corr_lookup_plan_set_lanes": (_i, [_vp, _i]),
    "rcb_corr_lookup_backward": (_i, [ctypes.POINTER(_vp), _vp, _vp, ctypes.POINTER(_vp), _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "rcb_corr_pool_backward": (_i, [ctypes.POINTER(_vp), _i, _i, _i, _i, _vp]),
    "rcb_corr_contract_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "rcb_corr_contract_backward_tc_workspace_bytes": (ctypes.c_size_t, [_i, _i, _i, _i]),
    "rcb_corr_contract_backward_tc": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, ctypes.c_size_t, _vp]),
    "rcb_altcorr_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "rcb_altcorr_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "rcb_altcorr_prepare": (_i, [_vp, _vp, _vp, ctypes.POINTER(_vp), _i, _i, _i, _i, _i, _vp]),
    "rcb_altcorr_pyramid_forward": (_i, [_vp, ctypes.POINTER(_vp), _vp, _vp, _i, _i, _i, _i, _i, _i, ctypes.c_float, _vp]),
    "rcb_upsample_flow": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "rcb_upsample_flow_backward_workspace_bytes": (ctypes.c_size_t, [_i, _i, _i]),
    "rcb_upsample_flow_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_size_t, _i, _i, _i, _vp]),
    "rcb_corr_convc1_pack_bytes": (ctypes.c_size_t, [_i, _i, _i]),
    "rcb_corr_convc1_pack": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "rcb_corr_lookup_convc1": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
}

_lib = None


def lib():
    """Loads the shared library (once).  Raises if it has not been built: no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m raft_optical_flow_b200.build` "
                "(nvcc, sm_100a).  raft_optical_flow_b200 has no CPU/PyTorch fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        variant = bool(os.environ.get("RCB_LIB_VARIANT"))  # A/B timing builds may be older revisions of the ABI
        for name, (res, args) in SIGNATURES.items():
            if variant and not hasattr(handle, name):
                continue
            fn = getattr(handle, name)  # AttributeError if the library does not export the ABI
            fn.restype = res
            fn.argtypes = args
        if handle.rcb_abi_version() != ABI_VERSION and not variant:
            raise RuntimeError("libraftcorr_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(status, what):
    if status != 0:
        msg = lib().rcb_status_string(status).decode()
        raise RuntimeError(f"{what} failed: {msg} (status {status})")


def ptr_array(ptrs):
    return (_vp * len(ptrs))(*[_vp(p) for p in ptrs])


def pyramid_layout(B, H, W, levels, dtype=F32):
    lay = PyramidLayout()
    check(lib().rcb_pyramid_layout_query(B, H, W, levels, dtype, ctypes.byref(lay)), "rcb_pyramid_layout_query")
    return lay


class LookupPlan:
    """Host-side lookup plan (rcb_corr_lookup_plan_*): the TMA tensor maps of one pyramid, encoded once."""

    def __init__(self, ptrs, B, H, W, levels, radius, dtype=F32):
        n = lib().rcb_corr_lookup_plan_bytes()
        self._raw = ctypes.create_string_buffer(n + 64)
        self.ptr = (ctypes.addressof(self._raw) + 63) & ~63
        self.nbytes = n
        check(lib().rcb_corr_lookup_plan_init(self.ptr, n, ptrs, B, H, W, levels, radius, dtype),
              "rcb_corr_lookup_plan_init")

    def set_lanes(self, lanes):
        """Pins the lanes per query of the fp32 lookup kernel (2 or 4; 0 = chosen per launch from the grid size)."""
        check(lib().rcb_corr_lookup_plan_set_lanes(self.ptr, lanes), "rcb_corr_lookup_plan_set_lanes")
