"""ctypes binding of include/mce_b200.h (libmce_b200.so).  No compute lives here: every call crosses the C ABI."""
import ctypes as ct
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# MCE_B200_LIB selects another build of the same library (A/B measurements of kernel variants; tools/ab_build.sh)
LIB_PATH = os.environ.get("MCE_B200_LIB") or os.path.join(PKG_DIR, "libmce_b200.so")
MAXM = 32


class MceOptions(ct.Structure):
    _fields_ = [("device", ct.c_int), ("tr_search_order", ct.c_int * 12), ("print_basic_info", ct.c_int),
                ("fast_moments", ct.c_int), ("group_split_threshold", ct.c_int), ("phase_timing", ct.c_int), ("lean_group_kernel", ct.c_int), ("early_scale_min_slots", ct.c_int), ("fast_moments_min_slots", ct.c_int), ("reserved", ct.c_int * 2)]


class MceMoments(ct.Structure):
    _fields_ = [("fz", ct.c_double * 2), ("fz_after_mu", ct.c_double * 2), ("mean", ct.c_double * 32), ("cov", ct.c_double * 512), ("mean_after_mu", ct.c_double * 32), ("cov_after_mu", ct.c_double * 512),
                ("g_scale_factor", ct.c_double), ("numeric_moment_errors", ct.c_int), ("Nt", ct.c_int), ("Nt_after_muc", ct.c_int),
                ("master_step", ct.c_int), ("skip_post_mu", ct.c_int)]


class MceStepStats(ct.Structure):
    _fields_ = [(n, ct.c_double) for n in ("ms_total", "ms_tp", "ms_mu", "ms_moments", "ms_regroup", "ms_ftr", "ms_gtable", "ms_compact")] + \
               [(n, ct.c_longlong) for n in ("parents", "slots", "terms_after_muc", "groups", "survivors", "bytes_gtable_algorithmic",
                                             "bytes_step_algorithmic", "kernel_launches")] + \
               [("ftr_rounds_max", ct.c_int), ("diag_unmodelled_alias", ct.c_int), ("diag_hash_overflow", ct.c_int),
                ("ev_step_ms", ct.c_double), ("ev_gtable_ms", ct.c_double), ("gtable_launches", ct.c_longlong),
                ("cells_parents", ct.c_longlong), ("cells_survivors", ct.c_longlong), ("split_groups", ct.c_longlong), ("ev_moments_ms", ct.c_double), ("ev_ftr_ms", ct.c_double), ("ev_mu_ms", ct.c_double), ("gtable_lean_launches", ct.c_longlong)]


class MceAllToAllV(ct.Structure):      # mce_alltoallv_args
    _fields_ = [("send", ct.c_void_p), ("recv", ct.c_void_p), ("soff", ct.POINTER(ct.c_longlong)), ("scnt", ct.POINTER(ct.c_longlong)),
                ("roff", ct.POINTER(ct.c_longlong)), ("rcnt", ct.POINTER(ct.c_longlong))]


class MceShardStats(ct.Structure):     # mce_shard_stats
    _fields_ = [("rank", ct.c_int), ("world", ct.c_int), ("owned_terms", ct.c_int), ("imported_parents", ct.c_int), ("local_parents", ct.c_int), ("pad_", ct.c_int),
                ("bytes_terms", ct.c_longlong), ("bytes_parents", ct.c_longlong), ("bytes_moments", ct.c_longlong), ("bytes_keys", ct.c_longlong), ("ms_stage", ct.c_double * 8)]


EXCHANGE_FN = ct.CFUNCTYPE(ct.c_int, ct.c_void_p, ct.c_int, ct.c_void_p, ct.c_longlong)

# every symbol include/mce_b200.h declares
SYMBOLS = ["mce_default_options", "mce_create", "mce_destroy", "mce_step", "mce_get_moments", "mce_shape_range",
           "mce_get_terms_per_shape", "mce_set_master_step", "mce_reset", "mce_reinitialize_start_statistics", "mce_set_first_term", "mce_shift_b",
           "mce_deterministic_time_prop", "mce_export_shape", "mce_cpdf_grid_count", "mce_marginal_1d_points", "mce_marginal_1d_grid", "mce_marginal_2d_points", "mce_marginal_2d_grid", "mce_cpdf_last_ms", "mce_get_step_stats", "mce_debug_div_selftest", "mce_debug_moment_sums", "mce_debug_sum_scan", "mce_debug_export_slots", "mce_debug_capture", "mce_debug_muc_shape", "mce_shard_unique_id", "mce_shard_init", "mce_shard_init_callback", "mce_shard_set_moments_mode", "mce_shard_export_gpos", "mce_shard_get_stats",
           "mce_last_error", "mce_version"]


def bind(lib):
    dp, ip = ct.POINTER(ct.c_double), ct.POINTER(ct.c_int)
    lib.mce_create.restype = ct.c_void_p
    lib.mce_create.argtypes = [ct.c_int] * 5 + [dp] * 5 + [ct.POINTER(MceOptions)]
    lib.mce_destroy.argtypes = [ct.c_void_p]
    lib.mce_destroy.restype = None
    lib.mce_step.restype = ct.c_int
    lib.mce_step.argtypes = [ct.c_void_p, ct.c_double, dp, dp, dp, dp, ct.c_double, dp, dp]
    lib.mce_get_moments.argtypes = [ct.c_void_p, ct.POINTER(MceMoments)]
    lib.mce_shape_range.argtypes = [ct.c_void_p]
    lib.mce_get_terms_per_shape.argtypes = [ct.c_void_p, ip, ct.c_int]
    lib.mce_set_master_step.argtypes = [ct.c_void_p, ct.c_int]
    lib.mce_set_master_step.restype = None
    lib.mce_reset.argtypes = [ct.c_void_p]
    lib.mce_reinitialize_start_statistics.argtypes = [ct.c_void_p, dp, dp, dp]
    lib.mce_set_first_term.argtypes = [ct.c_void_p, dp, dp, dp]
    lib.mce_shift_b.argtypes = [ct.c_void_p, dp, ct.c_double]
    lib.mce_deterministic_time_prop.argtypes = [ct.c_void_p, dp, dp, dp]
    lib.mce_export_shape.argtypes = [ct.c_void_p, ct.c_int, ip, ct.POINTER(ct.c_longlong), dp, dp, dp, ip, ct.POINTER(ct.c_uint32), dp]
    lib.mce_get_step_stats.argtypes = [ct.c_void_p, ct.POINTER(MceStepStats)]
    lib.mce_cpdf_grid_count.argtypes = [ct.c_double, ct.c_double, ct.c_double]
    lib.mce_marginal_1d_points.argtypes = [ct.c_void_p, ct.c_int, dp, ct.c_int, dp, dp]
    lib.mce_marginal_1d_grid.argtypes = [ct.c_void_p, ct.c_int, dp, ct.c_double, ct.c_double, ct.c_double, dp, ct.c_int]
    lib.mce_cpdf_last_ms.argtypes = [ct.c_void_p]
    lib.mce_marginal_2d_points.argtypes = [ct.c_void_p, ct.c_int, ct.c_int, dp, ct.c_int, dp, dp, dp]
    lib.mce_marginal_2d_grid.argtypes = [ct.c_void_p, ct.c_int, ct.c_int, dp] + [ct.c_double] * 6 + [dp, ct.c_int, ip, ip]
    lib.mce_cpdf_last_ms.restype = ct.c_double
    lib.mce_debug_capture.argtypes = [ct.c_void_p, ct.c_int]
    lib.mce_debug_sum_scan.argtypes = [ct.c_void_p, ct.c_longlong, dp, dp]
    lib.mce_debug_export_slots.argtypes = [ct.c_void_p, ct.c_longlong, dp, dp]
    lib.mce_debug_export_slots.restype = ct.c_longlong
    lib.mce_debug_moment_sums.argtypes = [ct.c_void_p, ct.c_longlong, ct.c_int, dp, dp, dp]
    lib.mce_debug_div_selftest.argtypes = [ct.c_void_p, ct.c_longlong, ct.c_ulonglong, ct.POINTER(ct.c_ulonglong)]
    lib.mce_shard_unique_id.argtypes = [ct.c_int, ct.c_void_p]
    lib.mce_shard_init.argtypes = [ct.c_void_p, ct.c_int, ct.c_int, ct.c_void_p]
    lib.mce_shard_init_callback.argtypes = [ct.c_void_p, ct.c_int, ct.c_int, EXCHANGE_FN, ct.c_void_p]
    lib.mce_shard_set_moments_mode.argtypes = [ct.c_void_p, ct.c_int]
    lib.mce_shard_export_gpos.argtypes = [ct.c_void_p, ip, ct.c_int]
    lib.mce_shard_get_stats.argtypes = [ct.c_void_p, ct.POINTER(MceShardStats)]
    lib.mce_debug_muc_shape.argtypes = [ct.c_void_p, ct.c_int, ip, dp, dp, dp, dp, dp, ip, ct.POINTER(ct.c_uint8), ct.POINTER(ct.c_int8), ip]
    lib.mce_last_error.restype = ct.c_char_p
    lib.mce_version.restype = ct.c_char_p
    lib.mce_default_options.argtypes = [ct.POINTER(MceOptions)]
    lib.mce_default_options.restype = None
    return lib


_LIB = None


def load():
    """Loads libmce_b200.so. Raises when the CUDA extension has not been built: there is no fallback."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("cauchyfriendly_b200: %s is missing; build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)" % LIB_PATH)
        _LIB = bind(ct.CDLL(LIB_PATH))
    return _LIB
