"""ctypes binding of libmultinn_sm100.so (the C ABI declared in include/multinn_b200.h).

There is no CPU fallback: importing this module without the built library, or calling an op with a
non-CUDA tensor, raises. Build with `make -C multinn_b200/csrc` (or `__graft_entry__.build()`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmultinn_sm100.so")


class MultinnLibraryError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise MultinnLibraryError(
            f"{LIB_PATH} is missing: the sm_100a CUDA library is the only compute path "
            "(run `make -C multinn_b200/csrc` or `python -c 'import __graft_entry__ as g; g.build()'`)")
    return C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)


lib = _load()

_p, _i, _ll, _f, _u64, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_ulonglong, C.c_size_t

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/multinn_b200.h
SIGNATURES = {
    "mnn_version": [],
    "mnn_last_error_string": [],
    "mnn_launch_count": [],
    "mnn_pack_pianoroll": [_p, _p, _p, _p, _i, _i, _i, _i, _p],
    "mnn_pack_pianoroll_u8": [_p, _p, _p, _p, _i, _i, _i, _i, _p],
    "mnn_pack_rows": [_p, _ll, _i, _p, _i, _i, _p],
    "mnn_gemm_f32": [_p, _ll, _i, _p, _ll, _i, _p, _ll, _p, _f, _f, _i, _i, _i, _p],
    "mnn_set_sm_budget": [_i],
    "mnn_set_gemm_split": [_i],
    "mnn_set_row_map": [_ll, _ll, _ll],
    "mnn_set_time_base": [_ll],
    "mnn_gemm_tc_supported": [_p, _ll, _p, _ll],
    "mnn_gemm_tc": [_p, _ll, _i, _p, _ll, _i, _p, _ll, _p, _f, _f, _i, _i, _i, _i, _p],
    "mnn_gemm_tc_bpair": [_p, _ll, _i, _p, _ll, _i, _p, _ll, _p, _f, _f, _i, _i, _i, _i, _p],
    "mnn_split_bf16_pair": [_p, _ll, _i, _i, _p, _ll, _p],
    "mnn_pack_stacked_bf16": [_p, _i, _p, _ll, _i, _i, _i, _p],
    "mnn_gemm_tc_abf16": [_p, _ll, _i, _p, _ll, _p, _ll, _i, _p, _ll, _p, _f, _f, _i, _i, _i, _p],
    "mnn_lstm_cell_fwd": [_p, _p, _p, _p, _p, _p, _p, _f, _u64, _u64, _i, _i, _p],
    "mnn_lstm_seq_fwd": [_p, _p, _p, _p, _p, _p, _p, _f, _u64, _i, _i, _i, _p],
    "mnn_lstm_seq_bwd": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p],
    "mnn_lstm_workspace_bytes": [_i, _i],
    "mnn_lstm_tc_supported": [_i, _i],
    "mnn_lstm_seq_fwd_ctas": [_i, _i, _i],
    "mnn_lstm_seq_bwd_ctas": [_i, _i, _i],
    "mnn_lstm_seq_fwd_tc": [_p, _p, _p, _p, _p, _p, _p, _f, _u64, _i, _i, _i, _p, _i, _p],
    "mnn_lstm_seq_bwd_tc": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _i, _p],
    "mnn_lstm_seq_bwd_tc_chunk": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _i, _i, _p],
    "mnn_colsum_workspace_bytes": [_i],
    "mnn_colsum": [_p, _ll, _i, _i, _p, _i, _p, _p],
    "mnn_nade_logprob_fwd": [_p, _p, _ll, _i, _i, _p, _p, _p, _p, _p, _f, _i, _i, _i, _i, _ll, _p],
    "mnn_set_nade_mode": [_i],
    "mnn_nade_logprob_bwd": [_p, _p, _ll, _i, _i, _p, _p, _p, _p, _p, _i, _i, _i, _i, _ll, _p],
    "mnn_nade_sample": [_p, _ll, _i, _i, _p, _p, _p, _i, _u64, _u64, _p, _ll, _i, _i, _p, _i, _i, _i, _i, _p],
    "mnn_generate_fused_workspace_bytes": [_i, _i, _i, _i, _i, _i, _i, _i],
    "mnn_generate_fused": [_i, _i, _p, _p, _p, _p, _i, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _ll, _p, _i, _u64, _u64, _p,
                           _ll, _ll, _i, _i, _i, _i, _i, _p, _p],
    "mnn_bias_sigmoid_sample": [_p, _ll, _p, _ll, _p, _ll, _i, _u64, _u64, _p, _ll, _p, _ll, _i, _i, _p],
    "mnn_rbm_gibbs_smem_bytes": [_i, _i],
    "mnn_rbm_gibbs": [_p, _ll, _p, _p, _ll, _p, _ll, _p, _p, _i, _u64, _u64, _p, _ll, _p, _ll, _p, _ll, _i, _i, _i, _i, _p],
    "mnn_sigmoid_bwd": [_p, _ll, _p, _ll, _p, _ll, _i, _i, _p],
    "mnn_rbm_free_energy": [_p, _ll, _p, _ll, _p, _ll, _p, _ll, _p, _i, _i, _i, _p],
    "mnn_reduce_workspace_bytes": [],
    "mnn_sum": [_p, _sz, _p, _p, _f, _i, _p],
    "mnn_sqnorm": [_p, _sz, _p, _p, _p],
    "mnn_clip_adam": [_p, _p, _p, _p, _sz, _p, _f, _f, _f, _f, _f, _f, _i, _p],
    "mnn_scale_rows": [_p, _ll, _i, _p, _ll, _i, _p],
    "mnn_axpy": [_p, _p, _f, _sz, _p],
    "mnn_clip_sgd": [_p, _p, _sz, _p, _f, _f, _f, _p],
    "mnn_probe_mufu": [_p, _i, _i, _i, _i, _p],
}
_RESTYPES = {"mnn_last_error_string": C.c_char_p, "mnn_launch_count": C.c_ulonglong, "mnn_reduce_workspace_bytes": C.c_size_t,
             "mnn_colsum_workspace_bytes": C.c_size_t, "mnn_lstm_workspace_bytes": C.c_size_t,
             "mnn_rbm_gibbs_smem_bytes": C.c_size_t, "mnn_generate_fused_workspace_bytes": C.c_size_t}

for _name, _args in SIGNATURES.items():
    _fn = getattr(lib, _name, None)
    if _fn is None:
        continue  # test_abi.py checks every symbol of the header is exported
    _fn.argtypes = _args
    _fn.restype = _RESTYPES.get(_name, C.c_int)


def last_error():
    s = lib.mnn_last_error_string()
    return s.decode() if s else ""


def check(rc, what):
    """Raise on a non-zero return code of a C-ABI call (negative: argument error, positive: cudaError_t)."""
    if rc != 0:
        raise MultinnLibraryError(f"{what} failed with code {rc}: {last_error()}")
