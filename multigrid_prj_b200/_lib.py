"""ctypes loader of libmgb200.so.  Fails loudly if the library is missing: no fallback path."""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(PKG, "lib", "libmgb200.so")


class MgbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mgb200 error {code}: {msg}")
        self.code = code


class GmgConfigStruct(C.Structure):
    # mirrors mgb_gmg_config in include/mgb200.h
    _fields_ = [("n", C.c_size_t), ("length", C.c_double), ("alpha", C.c_double),
                ("levels", C.c_int), ("smoother", C.c_int), ("pre_smoother", C.c_int),
                ("n_pre", C.c_int), ("nu", C.c_int), ("restriction", C.c_int),
                ("coarse_tol", C.c_double), ("coarse_maxit", C.c_int), ("device", C.c_int),
                ("rank", C.c_int), ("n_ranks", C.c_int), ("nccl_id", C.c_ubyte * 128),
                ("tail_max_width", C.c_int), ("use_graph", C.c_int), ("rb_fast_arith", C.c_int),
                ("rb_fused", C.c_int), ("fuse_correction", C.c_int), ("fuse_residual", C.c_int), ("fuse_prolong", C.c_int), ("defer_norm", C.c_int),
                ("jacobi_omega", C.c_double),
                ("cycle_type", C.c_int), ("nu_pre", C.c_int), ("fmg", C.c_int), ("reserved_cycle", C.c_int)]


class GmgStatsStruct(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("graph_launches", C.c_uint64),
                ("coarse_iters_total", C.c_uint64), ("cycles", C.c_uint64),
                ("bytes_algorithmic", C.c_double), ("reserved", C.c_int * 8)]


class AmgConfigStruct(C.Structure):
    # mirrors mgb_amg_config in include/mgb200.h
    _fields_ = [("levels", C.c_int), ("eps", C.c_double), ("smoother", C.c_int), ("pre_sweeps", C.c_int),
                ("coarse_sweeps", C.c_int), ("post_sweeps", C.c_int), ("exact_order", C.c_int), ("device", C.c_int),
                ("start_index", C.c_int64 * 16), ("hybrid_gs", C.c_int), ("shard_min_rows", C.c_int),
                ("jacobi_omega", C.c_double), ("tail_max_rows", C.c_int), ("cycle_graph", C.c_int),
                ("coop_sweeps", C.c_int), ("device_setup", C.c_int), ("coarse_smoother", C.c_int), ("p2p", C.c_int),
                ("reserved", C.c_int * 1)]


# every symbol include/mgb200.h declares: name -> (restype, argtypes)
_vp, _i, _d = C.c_void_p, C.c_int, C.c_double
_pd, _pi = C.POINTER(C.c_double), C.POINTER(C.c_int)
SYMBOLS = {
    "mgb_gmg_config_default": (None, [C.POINTER(GmgConfigStruct)]),
    "mgb_gmg_config_fast": (None, [C.POINTER(GmgConfigStruct)]),
    "mgb_gmg_create": (_i, [C.POINTER(GmgConfigStruct), C.POINTER(_vp)]),
    "mgb_gmg_destroy": (None, [_vp]),
    "mgb_gmg_partition": (_i, [C.c_size_t, _i, _i, _i, _i, _pi, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "mgb_gmg_pool_layout": (_i, [C.c_size_t, _i, _i, _i, _i, _i, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "mgb_gmg_level_width": (_i, [_vp, _i, C.POINTER(C.c_size_t)]),
    "mgb_gmg_level_rows": (_i, [_vp, _i, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "mgb_gmg_set_rhs": (_i, [_vp, _vp]),
    "mgb_gmg_set_rhs_test": (_i, [_vp, _i]),
    "mgb_gmg_set_u": (_i, [_vp, _vp]),
    "mgb_gmg_get_u": (_i, [_vp, _vp]),
    "mgb_gmg_set_level": (_i, [_vp, _i, _i, _vp]),
    "mgb_gmg_get_level": (_i, [_vp, _i, _i, _vp]),
    "mgb_gmg_smooth": (_i, [_vp, _i, _i, _i, _i, _i]),
    "mgb_gmg_residual": (_i, [_vp, _i, _i, _i, _i, _pd]),
    "mgb_gmg_sumsq": (_i, [_vp, _i, _i, _pd]),
    "mgb_gmg_restrict": (_i, [_vp]),
    "mgb_gmg_prolong": (_i, [_vp, _i]),
    "mgb_gmg_set_cycle": (_i, [_vp, _i, _i, _i, _d, _i]),
    "mgb_gmg_set_cycle_type": (_i, [_vp, _i, _i, _i]),
    "mgb_gmg_fmg": (_i, [_vp]),
    "mgb_gmg_krylov": (_i, [_vp, _i, _i, _d, _i, _vp, _pi]),
    "mgb_gmg_set_defer_norm": (_i, [_vp, _i]),
    "mgb_gmg_set_stream_impl": (_i, [_vp, _i]),
    "mgb_gmg_cycle": (_i, [_vp, _pd, _pi]),
    "mgb_gmg_fine_leg": (_i, [_vp, _pd]),
    "mgb_gmg_solve": (_i, [_vp, _d, _i, _i, _vp, _pi]),
    "mgb_gmg_iterate": (_i, [_vp, _d, _pd, _pd]),
    "mgb_gmg_run_cycles": (_i, [_vp, _i, _pd]),
    "mgb_gmg_checksum": (_i, [_vp, _i, _i, C.POINTER(C.c_uint64)]),
    "mgb_gmg_uses_p2p": (_i, [_vp]),
    "mgb_amg_uses_p2p": (_i, [_vp]),
    "mgb_gmg_get_stats": (_i, [_vp, C.POINTER(GmgStatsStruct)]),
    "mgb_gmg_reset_stats": (_i, [_vp]),
    "mgb_gmg_stream": (_vp, [_vp]),
    "mgb_gmg_sync": (_i, [_vp]),
    "mgb_amg_config_default": (None, [C.POINTER(AmgConfigStruct)]),
    "mgb_amg_config_fast": (None, [C.POINTER(AmgConfigStruct)]),
    "mgb_amg_config_device": (None, [C.POINTER(AmgConfigStruct)]),
    "mgb_amg_n_levels": (_i, [_vp]),
    "mgb_amg_create_from_csr": (_i, [C.POINTER(AmgConfigStruct), C.c_size_t, _vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    "mgb_amg_destroy": (None, [_vp]),
    "mgb_amg_create_sharded": (_i, [C.POINTER(AmgConfigStruct), C.c_size_t, _vp, _vp, _vp, _vp, _i, _i, _vp, C.POINTER(_vp)]),
    "mgb_amg_create_from_system": (_i, [C.POINTER(AmgConfigStruct), _vp, _i, _i, _vp, C.POINTER(_vp)]),
    "mgb_fem_assemble_p1": (_i, [C.c_size_t, _vp, _vp, _vp, C.c_size_t, _vp, _i, _i, C.POINTER(_vp)]),
    "mgb_fem_synthetic": (_i, [C.c_size_t, C.c_uint64, _i, C.POINTER(_vp)]),
    "mgb_fem_synthetic_mesh": (_i, [C.c_size_t, C.c_uint64, _vp, _vp, _vp, _vp]),
    "mgb_system_info": (_i, [_vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "mgb_system_get": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "mgb_system_destroy": (None, [_vp]),
    "mgb_amg_partition": (_i, [C.c_size_t, _i, _i, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "mgb_amg_level_rows": (_i, [_vp, _i, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), _pi]),
    "mgb_amg_halo_plan": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "mgb_amg_level_info": (_i, [_vp, _i] + [C.POINTER(C.c_size_t)] * 4 + [_pi, _pi]),
    "mgb_amg_get_matrix": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "mgb_amg_get_schedule": (_i, [_vp, _i, _i, _vp]),
    "mgb_amg_get_vector": (_i, [_vp, _i, _i, _vp]),
    "mgb_amg_set_vector": (_i, [_vp, _i, _i, _vp]),
    "mgb_amg_smooth": (_i, [_vp, _i, _i, _i]),
    "mgb_amg_restrict": (_i, [_vp, _i]),
    "mgb_amg_prolong": (_i, [_vp, _i]),
    "mgb_amg_residual": (_i, [_vp, _i, _pd]),
    "mgb_amg_apply": (_i, [_vp, _pd]),
    "mgb_amg_solve": (_i, [_vp, _d, _i, _i, _i, _i, _vp, _pi]),
    "mgb_csr_create": (_i, [C.c_size_t, C.c_size_t, _vp, _vp, _vp, C.POINTER(_vp)]),
    "mgb_csr_destroy": (None, [_vp]),
    "mgb_csr_info": (_i, [_vp] + [C.POINTER(C.c_size_t)] * 3),
    "mgb_csr_get": (_i, [_vp, _vp, _vp, _vp]),
    "mgb_amg_select_coarse_nodes": (_i, [_vp, _d, C.c_int64, _vp, C.POINTER(C.c_size_t)]),
    "mgb_amg_build_prolongation": (_i, [_vp, _d, _vp, C.POINTER(_vp)]),
    "mgb_amg_build_coarse_matrix": (_i, [_vp, _vp, C.POINTER(_vp)]),
    "mgb_amg_checksum": (_i, [_vp, _i, _i, C.POINTER(C.c_uint64)]),
    "mgb_amg_get_stats": (_i, [_vp, C.POINTER(GmgStatsStruct)]),
    "mgb_amg_reset_stats": (_i, [_vp]),
    "mgb_amg_stream": (_vp, [_vp]),
    "mgb_amg_sync": (_i, [_vp]),
    "mgb_nccl_unique_id": (_i, [C.POINTER(C.c_ubyte * 128)]),
    "mgb_last_error": (C.c_char_p, []),
    "mgb_version": (C.c_char_p, []),
    "mgb_device_count": (_i, []),
    "mgb_timer_create": (_i, [C.POINTER(_vp)]),
    "mgb_timer_destroy": (None, [_vp]),
    "mgb_timer_start": (_i, [_vp, _vp]),
    "mgb_timer_stop": (_i, [_vp, _vp]),
    "mgb_timer_elapsed_ms": (_i, [_vp, _pd]),
}

_lib = None


def lib_path():
    return _LIB


def load():
    """Loads libmgb200.so and binds every declared symbol.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise MgbError(-1, f"{_LIB} is not built: run `python -m multigrid_prj_b200.build` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(_LIB, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise MgbError(rc, load().mgb_last_error().decode())
