"""ctypes binding of libsocksb200.so (C ABI: include/socks_b200.h).

There is no fallback: if the shared library is missing, or no CUDA device is visible, the first
call raises.  Build with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C socks_b200/csrc`.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsocksb200.so")

_lib = None

_P, _I, _L, _D = c_void_p, c_int, c_int64, c_double

# name -> (restype, argtypes); must list every symbol include/socks_b200.h and include/socks_b200_dev.h declare
SIGNATURES = {
    "socks_version": (_I, []),
    "socks_last_error": (c_char_p, []),
    "socks_padded_dim": (_L, [_L]),
    "socks_copy2d": (_I, [_P, _L, _P, _L, _L, _L, _I, _P]),
    "socks_diag_range": (_I, [_P, _L, _L, _P, _P]),
    "socks_gram_rbf": (_I, [_P, _L, _P, _L, _I, _D, _I, _P, _P, _L, _P]),
    "socks_gram_kernel": (_I, [_I, _P, _L, _P, _L, _I, _D, _I, _P, _P, _L, _P]),
    "socks_gram_family": (_I, [_I, _P, _L, _P, _L, _I, _D, _I, _D, _D, _P, _P, _L, _P]),
    "socks_gram_sym_family": (_I, [_I, _P, _L, _I, _P, _I, _D, _I, _D, _D, _D, _P, _L, _P]),
    "socks_dist": (_I, [_P, _L, _P, _L, _I, _P, _L, _P]),
    "socks_gram_sym_kernel": (_I, [_I, _P, _L, _I, _D, _I, _D, _P, _L, _P]),
    "socks_colsumsq": (_I, [_P, _L, _L, _L, _P, _P, _P]),
    "socks_sqdist": (_I, [_P, _L, _P, _L, _I, _P, _L, _P]),
    "socks_gram_sym": (_I, [_P, _L, _I, _D, _I, _D, _P, _L, _P]),
    "socks_potrf_work_bytes": (_L, [_L]),
    "socks_trtri_work_bytes": (_L, [_L]),
    "socks_reduce_work_bytes": (_L, [_L]),
    "socks_potrf": (_I, [_P, _L, _L, _P, _P, _P]),
    "socks_pad_spd": (_I, [_P, _L, _L, _D, _P, _L, _P]),
    "socks_potrf_inv": (_I, [_P, _L, _L, _P, _L, _P, _P, _I, _P]),
    "socks_potrf_inv_set_block": (_I, [_I]),
    "socks_dev_potrf_trace": (_I, [_P, _P, _I]),
    "socks_trtri": (_I, [_P, _L, _L, _P, _P, _L, _P, _P]),
    "socks_diag_inv": (_I, [_P, _L, _L, _P, _P, _P]),
    "socks_potrs": (_I, [_P, _L, _L, _P, _L, _L, _P, _P]),
    "socks_lauum": (_I, [_P, _L, _L, _P, _L, _P]),
    "socks_lauum_i8_work_bytes": (_L, [_L]),
    "socks_lauum_i8": (_I, [_P, _L, _L, _P, _L, _P, _P]),
    "socks_sr_recursion": (_I, [_P, _L, _L, _L, _P, _I, _P, _I, _I, _P, _L, _P, _L, _P, _P]),
    "socks_sr_pack_bytes": (_L, [_L, _I, _I]),
    "socks_sr_pack": (_I, [_P, _P, _P, _L, _L, _I, _D, _I, _I, _P, _P]),
    "socks_sr_predict_packed_work_bytes": (_L, [_L, _L]),
    "socks_sr_predict_packed": (_I, [_P, _P, _L, _I, _D, _I, _P, _L, _P, _I, _I, _P, _L, _P, _I, _P]),
    "socks_sr_predict_work_bytes": (_L, [_L, _L]),
    "socks_sr_predict": (_I, [_P, _P, _P, _L, _L, _I, _D, _I, _P, _L, _P, _I, _I, _P, _L, _P, _I, _P]),
    "socks_srmax_pack": (_I, [_P, _L, _I, _P, _P, _L, _I, _I, _I, _P, _L, _L, _P]),
    "socks_gemm_tn": (_I, [_P, _L, _P, _L, _P, _L, _L, _L, _L, _P]),
    "socks_gemm_f64": (_I, [_I, _I, _P, _L, _P, _L, _P, _L, _L, _L, _L, _D, _D, _I, _I, _I, _P]),
    "socks_srmax_step": (_I, [_P, _L, _P, _L, _L, _I, _I, _I, _P, _I, _P, _I, _I, _P, _L, _I, _P]),
    "socks_scale_pack": (_I, [_P, _L, _I, _P, _L, _I, _P, _L, _L, _P]),
    "socks_rowmin_masked": (_I, [_P, _L, _L, _I, _I, _P, _P, _P]),
    "socks_gemv_t": (_I, [_P, _L, _L, _L, _P, _L, _I, _P, _L, _P, _P]),
    "socks_fwd_weights": (_I, [_P, _L, _I, _D, _I, _P, _P, _L, _I, _P, _L, _P]),
    "socks_argmin_masked": (_I, [_P, _P, _I, _P, _P]),
    "socks_atb_small_work_bytes": (_L, [_L, _I, _I]),
    "socks_atb_small": (_I, [_P, _L, _I, _P, _L, _I, _L, _P, _L, _P, _P]),
    "socks_linear_map": (_I, [_P, _I, _I, _P, _I, _I, _I, _P, _P, _L, _P, _L, _P]),
    "socks_sr_fit_work_bytes": (_L, [_L]),
    "socks_sr_fit": (_I, [_P, _P, _P, _L, _I, _D, _I, _P, _I, _I, _P, _L, _P, _P, _P]),
    "socks_srmax_fit_work_bytes": (_L, [_L, _I]),
    "socks_srmax_fit": (_I, [_P, _P, _P, _P, _P, _L, _I, _I, _I, _D, _I, _P, _I, _I, _P, _P, _L, _P, _P]),
    "socks_srmax_fit_i8_work_bytes": (_L, [_L, _I, _I]),
    "socks_srmax_fit_i8": (_I, [_P, _P, _P, _P, _P, _L, _I, _I, _I, _D, _I, _P, _I, _I, _P, _P, _L, _P, _I, _D, _P, _P]),
    "socks_srmax_predict_work_bytes": (_L, [_L, _I, _I, _L]),
    "socks_srmax_predict": (_I, [_P, _P, _P, _L, _P, _L, _I, _I, _D, _I, _P, _L, _P, _I, _I, _P, _L, _P, _L, _P]),
    "socks_srmax_predict_i8_work_bytes": (_L, [_L, _I, _I, _L, _I]),
    "socks_srmax_predict_i8": (_I, [_P, _P, _P, _L, _P, _L, _I, _I, _D, _I, _P, _L, _P, _I, _I, _P, _L, _P, _L, _I, _D, _P, _P]),
    "socks_fwd_scores_work_bytes": (_L, [_L, _I]),
    "socks_fwd_scores": (_I, [_P, _L, _P, _P, _L, _I, _I, _D, _I, _P, _P, _L, _I, _P, _P, _P]),
    "socks_fwd_policy": (_I, [_P, _L, _P, _P, _L, _I, _I, _D, _I, _P, _P, _L, _I, _P, _P, _P, _P]),
    "socks_median_work_bytes": (_L, [_L, _L]),
    "socks_median_sq_dist": (_I, [_P, _L, _P, _L, _I, _I, _P, _P, _P]),
    "socks_ctx_create": (_I, [_I, _P]),
    "socks_ctx_destroy": (_I, [_P]),
    "socks_ctx_stream": (_P, [_P]),
    "socks_ctx_workspace": (_I, [_P, _L, _P]),
    "socks_ctx_synchronize": (_I, [_P]),
    "socks_comm_unique_id": (_I, [_P]),
    "socks_comm_init": (_I, [_P, _I, _I, _P]),
    "socks_bcast": (_I, [_P, _P, _L, _I]),
    "socks_allgather": (_I, [_P, _P, _P, _L]),
    "socks_dev_potf2_phases": (_I, [_P, _L, _P, _P, _P, _P]),
    "socks_dev_i8gemm_check": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "socks_dev_ozaki_gemm_work_bytes": (_L, [_L, _L, _L]),
    "socks_dev_ozaki_gemm_nt": (_I, [_P, _L, _P, _L, _P, _L, _L, _L, _L, _I, _I, _P, _P]),
    "socks_probe_fp64": (_I, [_I, _I, _P, _I, _P]),
}


class SocksError(RuntimeError):
    pass


def load():
    """Load the shared library and declare signatures (no CUDA call is made here)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SocksError(
            f"{LIB_PATH} not found: build the CUDA library first (make -C socks_b200/csrc). "
            "socks_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def call(name, *args):
    """Call an int-returning entry point, raising SocksError on a non-zero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.socks_last_error()
        raise SocksError(f"{name} failed with status {rc}: {msg.decode() if msg else ''}")
    return rc
