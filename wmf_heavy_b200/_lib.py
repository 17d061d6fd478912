"""ctypes binding of ``libwmfb200.so`` (C ABI declared in ``include/wmf_b200.h``).

There is no CPU fallback: if the shared library is missing, or no CUDA device is usable, every
operator raises.  Build the library with ``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C wmf_heavy_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwmfb200.so")

# error codes (include/wmf_b200.h)
E_ARG, E_WORKSPACE, E_OUTLET_RANGE, E_OUTLET_MASK, E_CODE5, E_CYCLE, E_TOO_LARGE, E_CAPACITY = range(-1, -9, -1)
ENC_INTERNAL, ENC_KEYPAD = 0, 1
ACC_I32, ACC_I64, ACC_F64 = 0, 1, 2
ALGO_AUTO, ALGO_WALK, ALGO_SWEEP = 0, 1, 2


class WmfError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libwmfb200 error {code}: {msg}")
        self.code = code
        self.msg = msg


class Shia5Params(C.Structure):
    _fields_ = [("dt", C.c_float), ("retorno_gr", C.c_float), ("retorno_aq", C.c_float),
                ("calc_niter", C.c_int32), ("speed_type", C.c_int32 * 3), ("hspeed_given", C.c_int32),
                ("strict", C.c_int32)]


class Shia3Params(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "dt", "max_capilar", "max_gravita", "max_aquifer", "drena", "retorno_gr", "k_perc_cap_to_grav",
        "k_perc_grav_to_aqu", "k_baseflow", "max_infil_mm_h", "et_scale")] + [
        ("proportional", C.c_int32), ("pad", C.c_int32)]


_vp, _i64, _i32, _sz, _f = C.c_void_p, C.c_int64, C.c_int32, C.c_size_t, C.c_float
_int = C.c_int

# name -> (restype, argtypes); every symbol include/wmf_b200.h declares
SIGNATURES = {
    "wmf_last_error": (C.c_char_p, []),
    "wmf_version": (_int, []),
    "wmf_device_info": (_int, [_vp]),
    "wmf_d8_accum_workspace": (_sz, [_i64, _i64, _int, _int]),
    "wmf_d8_accum": (_int, [_vp, _vp, _vp, _int, _i64, _i64, _int, _int, _vp, _sz, _vp]),
    "wmf_d8_accum_tile_classes": (_int, [_vp, _sz, _i64, _i64, _vp, _vp]),
    "wmf_d8_stripe_workspace": (_sz, [_i64, _i64]),
    "wmf_d8_stripe_local": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _int, _vp, _vp, _sz, _vp]),
    "wmf_stripe_boundary_workspace": (_sz, [_int, _i64]),
    "wmf_stripe_boundary_inflow": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp, _sz, _vp]),
    "wmf_d8_stripe_finish": (_int, [_vp, _vp, _vp, _int, _i64, _i64, _i64, _i64, _int, _vp, _vp, _vp, _sz, _vp]),
    "wmf_forest_accum_workspace": (_sz, [_i64]),
    "wmf_forest_accum": (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "wmf_basin_mask_workspace": (_sz, [_i64, _i64]),
    "wmf_basin_mask": (_int, [_vp, _vp, _i64, _i64, _int, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "wmf_basin_structure_workspace": (_sz, [_i64, _i64]),
    "wmf_basin_structure": (_int, [_vp, _i64, _i64, _i32, _i32, _vp, _i64, _vp, _vp, _sz, _vp]),
    "wmf_coord2fil_col_host": (None, [_f, _f, _f, _f, _f, _i32, _i32, _vp, _vp]),
    "wmf_basin_vec_workspace": (_sz, [_i64]),
    "wmf_basin_acum_vec": (_int, [_vp, _i64, _i64, _vp, _vp, _sz, _vp]),
    "wmf_basin_basics": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "wmf_basin_coordxy": (_int, [_vp, _i64, _i64, _f, _f, _f, _i32, _vp, _vp, _vp]),
    "wmf_basin_map2basin": (_int, [_vp, _i64, _i64, _vp, _i64, _i64, _int, _vp, _vp]),
    "wmf_basin_2map": (_int, [_vp, _i64, _i64, _vp, _i64, _i64, _int, _vp, _vp]),
    "wmf_stream_nod": (_int, [_vp, _i64, _i64, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "wmf_stream_receiver_workspace": (_sz, [_i64, _i64]),
    "wmf_stream_receiver": (_int, [_vp, _vp, _vp, _i64, _i64, _int, C.c_double, C.c_double, _vp, _vp, _vp, _sz, _vp]),
    "wmf_hand_workspace": (_sz, [_i64, _i64]),
    "wmf_hand": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _int, _vp, _vp, _sz, _vp]),
    "wmf_path_sum_workspace": (_sz, [_i64, _i64]),
    "wmf_path_sum": (_int, [_vp, _vp, _vp, _i64, _i64, _int, C.c_double, C.c_double, _int, _vp, _vp, _sz, _vp]),
    "wmf_time_to_out_workspace": (_sz, [_i64, _i64]),
    "wmf_time_to_out": (_int, [_vp, _vp, _vp, _i64, _i64, _int, C.c_double, C.c_double, _vp, _vp, _sz, _vp]),
    "wmf_stream_nodes_workspace": (_sz, [_i64, _i64]),
    "wmf_stream_nodes": (_int, [_vp, _vp, _i64, _i64, _int, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "wmf_shia5_workspace": (_sz, [_i64, _i64, _i32, _i32]),
    "wmf_shia5_run": (_int, [_vp, _i64, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                             _vp, _vp, _sz, _vp]),
    "wmf_shia5_run_snap": (_int, [_vp, _i64, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                  _vp, _vp, _i32, _vp, _vp, _sz, _vp]),
    "wmf_shia5_run_ext": (_int, [_vp, _i64, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                 _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _sz, _vp]),
    "wmf_basin_depth_host": (_i32, [_vp, _i64, _vp]),
    "wmf_shia5_route_workspace": (_sz, [_i64]),
    "wmf_shia5_route_run": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp,
                                   _vp, _vp, _sz, _vp]),
    "wmf_fitness_nash_peak": (_int, [_vp, _vp, _i32, _i64, _vp, _vp]),
    "wmf_shia3_workspace": (_sz, [_i64, _i64]),
    "wmf_shia3_run": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                             _vp, _vp, _sz, _vp]),
    "wmf_dir_reclass_scan": (_int, [_vp, _int, _i64, _vp, _vp, _vp]),
    "wmf_dir_reclass_apply": (_int, [_vp, _int, _i64, _vp, _int, _i64, _vp, _vp]),
    "wmf_index_scan": (_int, [_vp, _i64, _i64, _vp, _vp]),
    "wmf_index_repair": (_int, [_vp, _i64, _i64, _vp]),
    "wmf_scatter_last": (_int, [_vp, _int, _vp, _i64, _vp, _vp, _i64, _vp]),
    "wmf_gather": (_int, [_vp, _int, _vp, _i64, _vp, _vp]),
    "wmf_synth_scheidegger": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, C.c_uint64, _int, _vp]),
    "wmf_synth_scheidegger_host": (None, [_vp, _i64, _i64, _i64, _i64, _i64, C.c_uint64, _int]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raise (never fall back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: the CUDA library has not been built "
                "(run __graft_entry__.build() or make -C wmf_heavy_b200/csrc). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().wmf_last_error().decode("utf-8", "replace")
        raise WmfError(rc, msg)
