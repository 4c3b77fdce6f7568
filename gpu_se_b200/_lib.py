"""ctypes binding of libgse_b200.so (the C ABI declared in include/gse.h).

There is no fallback: if the shared library is missing or cannot be loaded this module raises,
and every filter operation raises with the library's own error message on a non-zero status.
Build the library with ``python gpu_se_b200/build.py`` (``__graft_entry__.build()`` does that).
"""
import ctypes
import os

GSE_NX, GSE_NU, GSE_NY, GSE_NSIGMA, GSE_NCOV, GSE_MAX_ND = 5, 2, 2, 11, 15, 8
GSE_MODEL_BIOREACTOR = 1
GSE_ABI_VERSION = 7

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgse_b200.so")

c_i64, c_u64, c_int, c_dbl, c_vp = (ctypes.c_int64, ctypes.c_uint64, ctypes.c_int, ctypes.c_double,
                                    ctypes.c_void_p)
c_dbl_p = ctypes.POINTER(ctypes.c_double)


class GseError(RuntimeError):
    pass


class gse_mixture(ctypes.Structure):
    _fields_ = [("nd", ctypes.c_int32), ("nx", ctypes.c_int32),
                ("weights", ctypes.c_double * GSE_MAX_ND),
                ("means", ctypes.c_double * (GSE_MAX_ND * GSE_NX)),
                ("covs", ctypes.c_double * (GSE_MAX_ND * GSE_NX * GSE_NX))]


GSE_ERR_CHOLESKY, GSE_ERR_SINGULAR_PYY, GSE_ERR_PEER_TIMEOUT, GSE_ERR_QUEUE_OVERFLOW, GSE_ERR_ZERO_WEIGHTS = 1, 2, 4, 8, 16
GSE_MAX_SHARDS, GSE_IPC_HANDLE_BYTES = 8, 64
GSE_MAILBOX_BYTES = 2 * GSE_MAX_SHARDS * 512


class gse_shards(ctypes.Structure):
    _fields_ = [("nshards", ctypes.c_int32), ("rows", ctypes.c_int64 * (GSE_MAX_SHARDS + 1)),
                ("cumsum_dev", ctypes.c_void_p * GSE_MAX_SHARDS), ("state_dev", ctypes.c_void_p * GSE_MAX_SHARDS),
                ("ld", ctypes.c_int64 * GSE_MAX_SHARDS), ("offsets_dev", ctypes.c_void_p),
                ("idx_dev", ctypes.c_void_p * GSE_MAX_SHARDS), ("rank", ctypes.c_int32)]


class gse_step_params(ctypes.Structure):
    _fields_ = [("u", ctypes.c_double * 2), ("dt", ctypes.c_double), ("z", ctypes.c_double * 2),
                ("r", ctypes.c_double), ("step", ctypes.c_uint64), ("reserved", ctypes.c_uint64)]


c_mix_p = ctypes.POINTER(gse_mixture)
c_shards_p = ctypes.POINTER(gse_shards)

# name -> (restype, argtypes); mirrors include/gse.h one to one (tests/test_abi.py checks the two
# against each other)
SIGNATURES = {
    "gse_abi_version": (c_int, []),
    "gse_last_error": (ctypes.c_char_p, []),
    "gse_ctx_create": (c_int, [c_int, c_int, c_i64, c_mix_p, c_mix_p, ctypes.POINTER(c_vp)]),
    "gse_ctx_destroy": (c_int, [c_vp]),
    "gse_mixture_draw": (c_int, [c_vp, c_mix_p, c_vp, c_i64, c_i64, c_u64, c_u64, c_i64, c_vp]),
    "gse_mixture_pdf": (c_int, [c_vp, c_mix_p, c_vp, c_i64, c_i64, c_vp, c_int, c_vp]),
    "gse_pf_predict": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_dbl_p, c_dbl, c_int, c_u64, c_u64, c_i64,
                               c_vp, c_i64, c_vp]),
    "gse_pf_update": (c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_dbl_p, c_dbl_p, c_vp, c_vp]),
    "gse_pf_moments": (c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_int, c_vp, c_vp]),
    "gse_loglik_max": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "gse_weights_linear": (c_int, [c_vp, c_vp, c_vp, c_i64, c_dbl, c_vp, c_vp]),
    "gse_scan_weights": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "gse_resample_search": (c_int, [c_vp, c_vp, c_i64, c_vp, c_dbl, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "gse_resample_fused": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_dbl, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_i64,
                                   c_vp, c_int, c_vp]),
    "gse_resample_fused_sharded": (c_int, [c_vp, c_vp, c_vp, c_vp, c_dbl, c_shards_p, ctypes.POINTER(c_vp), c_int,
                                           ctypes.c_uint, ctypes.c_uint, c_vp, c_vp, c_i64, c_vp, c_int, c_vp]),
    "gse_resample_search_f64": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_dbl, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "gse_ctx_errors": (ctypes.c_uint, [c_vp, c_int]),
    "gse_ctx_result_block": (c_int, [c_vp, ctypes.POINTER(c_dbl_p), ctypes.POINTER(c_dbl_p)]),
    "gse_ctx_wait": (c_int, [c_vp, c_vp, ctypes.POINTER(ctypes.c_uint)]),
    "gse_gather_rows": (c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_int, c_vp, c_vp]),
    "gse_peer_alloc": (c_int, [c_int, c_i64, ctypes.POINTER(c_vp), ctypes.c_char_p]),
    "gse_peer_free": (c_int, [c_int, c_vp]),
    "gse_peer_open": (c_int, [c_int, ctypes.c_char_p, ctypes.POINTER(c_vp)]),
    "gse_peer_close": (c_int, [c_int, c_vp]),
    "gse_resample_search_sharded": (c_int, [c_vp, c_shards_p, c_dbl, c_i64, c_i64, c_vp, c_vp]),
    "gse_gather_rows_sharded": (c_int, [c_vp, c_shards_p, c_vp, c_i64, c_vp, c_i64, c_int, c_vp]),
    "gse_pf_can_fuse_update": (c_int, [c_vp, c_int, c_i64]),
    "gse_pf_predict_update": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_dbl_p, c_dbl, c_int, c_u64, c_u64,
                                      c_i64, c_dbl_p, c_vp, c_vp, c_vp, c_vp]),
    "gse_pf_predict_update_sharded": (c_int, [c_vp, c_shards_p, c_vp, c_vp, c_i64, c_i64, c_dbl_p, c_dbl, c_int, c_u64,
                                              c_u64, c_i64, c_dbl_p, c_vp, c_vp, c_vp, c_vp]),
    "gse_pf_predict_sharded": (c_int, [c_vp, c_shards_p, c_vp, c_vp, c_i64, c_i64, c_dbl_p, c_dbl, c_int, c_u64, c_u64,
                                       c_i64, c_vp, c_i64, c_vp]),
    "gse_pf_moments_sharded": (c_int, [c_vp, c_shards_p, c_vp, c_i64, c_vp, c_vp, c_vp, c_int, c_vp, c_vp]),
    "gse_peer_allgather_stats": (c_int, [c_vp, ctypes.POINTER(c_vp), c_int, c_int, ctypes.c_uint, c_vp, c_vp]),
    "gse_peer_allgather_totals": (c_int, [c_vp, ctypes.POINTER(c_vp), c_int, c_int, ctypes.c_uint, c_vp, c_vp, c_vp]),
    "gse_peer_allgather_moments": (c_int, [c_vp, ctypes.POINTER(c_vp), c_int, c_int, ctypes.c_uint, c_vp, c_vp, c_vp,
                                           c_vp]),
    "gse_ctx_upload_step_params": (c_int, [c_vp, ctypes.POINTER(gse_step_params), c_vp]),
    "gse_ctx_use_step_params": (c_int, [c_vp, c_int]),
    "gse_merge_stats": (c_int, [c_vp, c_vp, c_int, c_vp, c_vp]),
    "gse_count_outputs_below": (c_i64, [c_u64, c_u64, c_dbl, c_i64]),
    "gse_threshold_u64": (c_u64, [c_dbl, c_u64]),
    "gse_gsf_predict": (c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, c_dbl_p, c_dbl, c_u64, c_u64,
                                c_i64, c_vp, c_i64, c_vp]),
    "gse_gsf_predict_sharded": (c_int, [c_vp, c_shards_p, c_vp, c_vp, c_vp, c_i64, c_i64, c_dbl_p, c_dbl, c_u64, c_u64,
                                        c_i64, c_vp, c_i64, c_vp]),
    "gse_gsf_moments_sharded": (c_int, [c_vp, c_shards_p, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gse_gsf_update": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_dbl_p, c_dbl_p, c_vp, c_vp]),
    "gse_gsf_sigma_points": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp]),
    "gse_gsf_moments": (c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "gse_ctx_read_trace": (c_i64, [c_vp, c_vp, c_i64]),
    "gse_launch_count": (c_i64, [c_vp]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise GseError("libgse_b200.so not found at %s -- build it first (python gpu_se_b200/build.py); "
                       "there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)           # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.gse_abi_version() != GSE_ABI_VERSION:
        raise GseError("libgse_b200.so ABI version %d, bindings expect %d" % (lib.gse_abi_version(), GSE_ABI_VERSION))
    return lib


lib = _load()


def check(status):
    """Raise GseError with the library's message on a non-zero status."""
    if status != 0:
        raise GseError("libgse_b200 call failed (status %d): %s" % (status, lib.gse_last_error().decode()))


_double2 = ctypes.c_double * 2


def as_double2(v):
    try:                                  # the usual case -- a length-2 vector -- without a trip through numpy
        if len(v) == 2 and getattr(v, "ndim", 1) == 1:
            return _double2(float(v[0]), float(v[1]))
    except TypeError:
        pass
    import numpy
    a = numpy.asarray(v, dtype=numpy.float64).ravel()
    if a.size != 2:
        raise ValueError("expected 2 values, got %d" % a.size)
    return _double2(float(a[0]), float(a[1]))


def make_mixture(means, covariances, weights):
    """gse_mixture from (nd, nx) means, (nd, nx, nx) covariances and (nd,) weights."""
    import numpy
    means = numpy.asarray(means, dtype=numpy.float64)
    covariances = numpy.asarray(covariances, dtype=numpy.float64)
    weights = numpy.asarray(weights, dtype=numpy.float64).ravel()
    if means.ndim != 2 or covariances.shape != (means.shape[0], means.shape[1], means.shape[1]) \
            or weights.shape != (means.shape[0],):
        raise ValueError("mixture shapes: means (nd, nx), covariances (nd, nx, nx), weights (nd,)")
    nd, nx = means.shape
    if nd > GSE_MAX_ND or nx > GSE_NX:
        raise ValueError("at most %d components of dimension <= %d are supported" % (GSE_MAX_ND, GSE_NX))
    m = gse_mixture()
    m.nd, m.nx = nd, nx
    for i, v in enumerate(weights):
        m.weights[i] = float(v)
    for i, v in enumerate(means.ravel()):
        m.means[i] = float(v)
    for i, v in enumerate(covariances.ravel()):
        m.covs[i] = float(v)
    return m
