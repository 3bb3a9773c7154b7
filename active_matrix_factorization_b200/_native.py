"""ctypes binding of libamf_b200.so (the C ABI in include/amf_b200.h).

There is no CPU implementation behind this module: if the shared library is missing or no
sm_100-class device is visible, importing callers get a loud RuntimeError.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# AMF_B200_LIB: an alternative build of the same library (kernel-variant timing, benchmarks/variant_lib.sh)
LIB_PATH = os.environ.get("AMF_B200_LIB") or os.path.join(HERE, "csrc", "libamf_b200.so")

F32, F64 = 0, 1
CRIT_PRED, CRIT_APPROX_MEAN, CRIT_PRED_VARIANCE, CRIT_PROB_GE = 0, 1, 2, 3


class PmfParams(C.Structure):
    _fields_ = [("sigma_sq", C.c_double), ("sigma_u_sq", C.c_double),
                ("sigma_v_sq", C.c_double), ("mean_offset", C.c_double)]


class NormalView(C.Structure):
    _fields_ = [("mean_u", C.c_void_p), ("mean_u_stride", C.c_int64),
                ("mean_v", C.c_void_p), ("mean_v_stride", C.c_int64),
                ("cov_uu", C.c_void_p), ("uu_stride", C.c_int64), ("uu_ld", C.c_int64),
                ("cov_vv", C.c_void_p), ("vv_stride", C.c_int64), ("vv_ld", C.c_int64),
                ("cov_uv", C.c_void_p), ("uv_stride_i", C.c_int64),
                ("uv_stride_j", C.c_int64), ("uv_ld", C.c_int64)]


class NormalFitParams(C.Structure):
    _fields_ = [("n", C.c_int32), ("m", C.c_int32), ("d", C.c_int32),
                ("sigma_sq", C.c_double), ("sigma_u_sq", C.c_double), ("sigma_v_sq", C.c_double),
                ("learning_rate", C.c_double), ("min_eig", C.c_double), ("kl_stop", C.c_double),
                ("min_lr", C.c_double), ("max_steps", C.c_int32)]


NORMAL_FIT, NORMAL_KL, NORMAL_GRADIENT, NORMAL_PROJECT = 0, 1, 2, 3


class BlocksView(C.Structure):
    _fields_ = [("n", C.c_int32), ("m", C.c_int32), ("d", C.c_int32),
                ("mean_u", C.c_void_p), ("cov_u", C.c_void_p), ("prec_u", C.c_void_p),
                ("h_u", C.c_void_p), ("logdet_u", C.c_void_p),
                ("mean_v", C.c_void_p), ("cov_v", C.c_void_p), ("prec_v", C.c_void_p),
                ("h_v", C.c_void_p), ("logdet_v", C.c_void_p),
                ("sums", C.c_void_p), ("sigma_sq", C.c_double), ("entropy0", C.c_double)]


LOOK_ENTROPY, LOOK_TOTAL_VARIANCE = 0, 1
WEIGHTS_NONE, WEIGHTS_DISCRETE, WEIGHTS_NODES = 0, 1, 2


class Best(C.Structure):
    _fields_ = [("value", C.c_double), ("index", C.c_int64)]


_P = C.c_void_p
_I32, _I64, _F64 = C.c_int32, C.c_int64, C.c_double
_INT = C.c_int

# name -> argtypes; every symbol declared in include/amf_b200.h appears here
PROTOTYPES = {
    "amf_version": [],
    "amf_device_info": [C.POINTER(_INT), C.POINTER(_INT), C.POINTER(_INT)],
    "amf_ratings_create": [C.POINTER(_P), _I32, _I32, _I64, _P, _P, _P, _INT, _P],
    "amf_ratings_create_host": [C.POINTER(_P), _I32, _I32, _I64, _P, _P, _P, _INT],
    "amf_ratings_destroy": [_P],
    "amf_ratings_nnz": [_P],
    "amf_ratings_layout": [_P, _INT, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)],
    "amf_ratings_mean": [_P, C.POINTER(_F64), _P],
    "amf_pmf_loss_grad": [_P, _INT, _INT, _INT, _P, _P, C.POINTER(PmfParams), _P, _P, _P, _P],
    "amf_pmf_loss_grad_part": [_P, _INT, _INT, _INT, _P, _P, C.POINTER(PmfParams), _P, _P, _P, _INT,
                               _INT, _P],
    "amf_axpy": [_INT, _I64, _P, _P, _F64, _P, _P],
    "amf_pmf_grad_coo": [_INT, _I64, _P, _P, _P, _INT, _INT, _P, _P, C.POINTER(PmfParams),
                         _P, _P, _P, _P],
    "amf_momentum_step": [_INT, _I64, _P, _P, _F64, _F64, _P, _P],
    "amf_pmf_prior": [_INT, _I64, _P, _F64, _P, _P, _P],
    "amf_pmf_loss_grad_host": [_P, _INT, _INT, _P, _P, C.POINTER(PmfParams), _P, _P, _P],
    "amf_score_candidates": [_INT, _INT, _I64, _P, _P, _INT, _INT, _P, _P,
                             C.POINTER(NormalView), _F64, _P, _INT, _I64, _P, _P],
    "amf_gibbs_half_sweep": [_P, _INT, _INT, _INT, _P, _P, _P, _F64, _F64, _P, _P, _P],
    "amf_gibbs_half_sweep_rows": [_P, _INT, _INT, _INT, _P, _P, _P, _F64, _F64, _P, _P, _I32, _I32, _P],
    "amf_gibbs_half_sweep_device_rng": [_P, _INT, _INT, _INT, _P, _P, _P, _F64, _F64, C.c_uint64,
                                        C.c_uint64, _P, _I32, _I32, _P],
    "amf_gibbs_half_sweep_batched": [_P, _INT, _INT, _INT, _INT, _P, _P, _P, _F64, _F64, _P, _P, _P, _P,
                                     C.c_uint64, C.c_uint64, _P, _P],
    "amf_gibbs_chain_device": [_P, _INT, _INT, _INT, _INT, _P, _P, _P, _P, _F64, _F64, C.c_uint64,
                               C.c_uint64, _P, _P, _P],
    "amf_gibbs_hyper_device": [_P, _INT, _INT, _I64, _P, _P, C.c_uint64, C.c_uint64, _P, _P, _P],
    "amf_philox_normal": [_INT, C.c_uint64, C.c_uint64, _I64, _INT, _P, _P],
    "amf_gibbs_status": [_P, C.POINTER(_INT), _P],
    "amf_bayes_sample_stats": [_INT, _I64, _P, _P, _INT, _I32, _I32, _INT, _P, _P, _F64, _F64,
                               _P, _P, _P, _INT, _INT, _I64, _P, _P],
    "amf_bayes_sample_stats_dense_tc": [_INT, _I32, _I32, _INT, _P, _P, _F64, _P, _P, _INT, _INT,
                                        _I64, _P, _P],
    "amf_normal_workspace_doubles": [_I32, _I32, _INT],
    "amf_normal_batched": [_INT, _INT, _I64, _P, _P, _P, _P, _P, _P, C.POINTER(NormalFitParams),
                           _P, _P, _P, _P, _P, _P, _INT, _P, _P, _P],
    "amf_pmf_fit_workspace_bytes": [_P, _INT, _INT],
    "amf_pmf_fit_lls": [_P, _INT, _INT, _INT, _P, _P, C.POINTER(PmfParams), _F64, _F64, _F64, _INT, _P,
                        _INT, _P, _P, _I64, _P],
    "amf_ratings_set_layout": [_P, _INT],
    "amf_ratings_append": [_P, _I64, _P, _P, _P, _P],
    "amf_ratings_compact": [_P, _P],
    "amf_best_reduce": [_P, _INT, _INT, _P, _P],
    "amf_peer_create": [C.POINTER(_P), _INT, _INT, _P],
    "amf_peer_connect": [_P, _P],
    "amf_peer_best_reduce": [_P, _P, _INT, _P, _P],
    "amf_peer_destroy": [_P],
    "amf_pred_covs": [_I32, _I32, _INT, _INT, _P, _P, _P, _P],
    "amf_slogdet_batched": [_INT, _INT, _P, _P, _P, _P],
    "amf_predicted_matrix": [_INT, _I32, _I32, _INT, _INT, _P, _P, _F64, _P, _P],
    "amf_sq_error_dense": [_INT, _I32, _I32, _INT, _INT, _P, _P, _F64, _P, _P, _P, _P],
    "amf_pool_max_tile_rows": [_INT],
    "amf_pool_create": [C.POINTER(_P), _I64, _P, _P, _I32, _I32, _INT, _P],
    "amf_pool_destroy": [_P],
    "amf_pool_size": [_P],
    "amf_pool_remove": [_P, _I64, _P, _P],
    "amf_pool_score_pred": [_P, _INT, _INT, _INT, _P, _P, _P, _INT, _I64, _P, _P],
    "amf_pool_score_pred_peer": [_P, _INT, _INT, _INT, _P, _P, _P, _INT, _I64, _P, _P, _P],
    "amf_mn_workspace_doubles": [_I32, _I32, _INT],
    "amf_mn_batched": [_INT, _INT, _I64, _P, _P, _P, _P, _P, _P, C.POINTER(NormalFitParams),
                       _P, _P, _P, _P, _P, _P, _P, _INT, _P, _P, _P],
    "amf_mn_score_candidates": [_INT, _INT, _I64, _P, _P, _I32, _I32, _INT, _P, _P, _P, _F64, _P,
                                _INT, _I64, _P, _P],
    "amf_blocks_half_sweep": [_P, _INT, _INT, _P, _P, _F64, _F64, _F64, _P, _P, _P, _P, _P, _P, _P],
    "amf_blocks_fit": [_P, _INT, _F64, _F64, _F64, _F64, _INT, _INT, _F64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                       _P, C.POINTER(_INT), _P],
    "amf_blocks_sums": [_I64, _INT, _P, _P, _P, _P],
    "amf_blocks_lookahead": [C.POINTER(BlocksView), _INT, _INT, _I64, _P, _P, _INT, _P, _INT, _P,
                             _P, _P, _P, _P, _INT, _I64, _P, _P, _P],
    "amf_blocks_pack": [_INT, _I64, _INT, _P, _P, _INT, _INT, _P, _P],
    "amf_prob_ge": [_INT, _I64, _P, _P, _F64, _P, _INT, _I64, _P, _P],
    "amf_score_pred_host": [_INT, _I64, _P, _P, _I32, _I32, _INT, _P, _P, _P, _INT,
                            C.POINTER(Best)],
    "amf_score_pred_host_csr": [_INT, _P, _P, _I32, _I32, _INT, _P, _P, _P, _INT, C.POINTER(Best)],
    "amf_score_pred_host_csr16": [_INT, _P, _P, _I32, _I32, _INT, _P, _P, _P, _INT, C.POINTER(Best)],
}

_lib = None


def library_path():
    return LIB_PATH


def load():
    """Loads the shared library (no device needed) and applies the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libamf_b200.so is not built (%s). Run `python -m active_matrix_factorization_b200.build`. "
            "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.amf_last_error.restype = C.c_char_p
    lib.amf_last_error.argtypes = []
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.argtypes = argtypes
        fn.restype = _I64 if name in ("amf_ratings_nnz", "amf_normal_workspace_doubles", "amf_pool_size",
                                    "amf_pmf_fit_workspace_bytes",
                                   "amf_mn_workspace_doubles") else _INT
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError("libamf_b200: " + load().amf_last_error().decode("utf-8", "replace"))


_device_checked = False


def require_device():
    """Fails loudly unless a CUDA device is usable.  Called by every compute path."""
    global _device_checked
    lib = load()
    if _device_checked:
        return lib
    n, arch, sms = _INT(), _INT(), _INT()
    check(lib.amf_device_info(C.byref(n), C.byref(arch), C.byref(sms)))
    if arch.value < 100:
        raise RuntimeError("libamf_b200 targets sm_100a; found sm_%d" % arch.value)
    _device_checked = True
    return lib


def dtype_code(np_dtype):
    np_dtype = np.dtype(np_dtype)
    if np_dtype == np.float32:
        return F32
    if np_dtype == np.float64:
        return F64
    raise TypeError("unsupported dtype %r" % (np_dtype,))


def host_ptr(arr):
    return arr.ctypes.data_as(C.c_void_p)
