"""ctypes binding of libfreud_b200.so (the C ABI declared in include/freud_b200.h).

There is no CPU fallback: if the library is missing or a call fails this raises.  The reference has no FFI
(its hot path is ATen ops); see INTEGRATION.md for how these entry points map onto its call sites.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfreud_b200.so")

MAX_TENSORS = 8
BF16, FP32 = 0, 1


class TensorList(C.Structure):
    _fields_ = [
        ("count", C.c_int32),
        ("param", C.c_void_p * MAX_TENSORS),
        ("grad", C.c_void_p * MAX_TENSORS),
        ("exp_avg", C.c_void_p * MAX_TENSORS),
        ("exp_avg_sq", C.c_void_p * MAX_TENSORS),
        ("bf16_shadow", C.c_void_p * MAX_TENSORS),
        ("numel", C.c_int64 * MAX_TENSORS),
    ]


_p, _i64, _i, _f, _d = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_double

# name -> argtypes, exactly the prototypes of include/freud_b200.h
SIGNATURES = {
    "freud_topk_prep_x": [_p, _i, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i, _p],
    "freud_split_operand": [_p, _p, _p, _i64, _i, _p],
    "freud_topk_encode_workspace": [_i64, _i64, _p],
    "freud_topk_encode": [_p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i, _p, _i64, _p, _p],
    "freud_topk_encode_stats": [_p, _i],
    "freud_gemm_nt": [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i, _i, _p],
    "freud_row_topk": [_p, _p, _p, _p, _i64, _i64, _i64, _p],
    "freud_gather_rows": [_p, _p, _p, _i64, _i64, _p],
    "freud_index_map": [_p, _p, _p, _i64, _p],
    "freud_row_topk_mask": [_p, _p, _i64, _i64, _i64, _i64, _i64, _i, _p],
    "freud_col_sum_bf16": [_p, _p, _i64, _i64, _i64, _p],
    "freud_scatter_add_rows": [_p, _p, _p, _i64, _i64, _p],
    "freud_sum_splits": [_p, _p, _i64, _i64, _p],
    "freud_gemm_nt_mask": [_p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _p],
    "freud_l1_encode_fused": [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _p],
    "freud_l1_decode_fused": [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _p],
    "freud_gemm_tn_splitk": [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _p],
    "freud_gemm_nn": [_p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i, _p],
    "freud_topk_decode": [_p, _p, _p, _i, _p, _p, _p, _p, _i, _p, _p, _i64, _i64, _i64, _p],
    "freud_topk_dacts": [_p, _i, _p, _p, _i, _p, _i64, _i64, _i64, _p],
    "freud_topk_decode_dacts_supported": [_i64, _i64],
    "freud_topk_decode_dacts": [_p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _i64, _i64, _i64, _p],
    "freud_topk_refine": [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _p],
    "freud_axpby": [_p, _p, _p, _p, _i, _i64, _p],
    "freud_shard_merge": [_p, _p, _p, _p, _i64, _i64, _i64, _p],
    "freud_shard_localize": [_p, _p, _p, _p, _i64, _i64, _i64, _p],
    "freud_residual": [_p, _p, _p, _i, _p, _p, _i64, _i64, _p],
    "freud_csc_build": [_p, _i64, _i64, _i64, _p, _p, _p, _i, _p],
    "freud_csc_meta": [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _p],
    "freud_topk_sparse_grads": [_p, _p, _p, _i, _p, _i, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i, _p],
    "freud_topk_bdec_grad": [_p, _p, _p, _p, _i, _p, _i64, _i64, _i, _p],
    "freud_topk_loss_scalars": [_p, _p, _p, _i64, _p],
    "freud_dead_latent_update": [_p, _p, _i64, _i64, _p],
    "freud_rownorm_project": [_p, _i64, _i64, _f, _p],
    "freud_remove_parallel_grad": [_p, _p, _i64, _i64, _p],
    "freud_l1_colnorm": [_p, _p, _i64, _i64, _p],
    "freud_l1_loss_scalars": [_p, _d, _d, _d, _p, _p],
    "freud_l1_loss_reduce": [_p, _p, _p, _p, _p, _i64, _i64, _i64, _p],
    "freud_l1_dz": [_p, _p, _p, _p, _i64, _i64, _p],
    "freud_l1_weight_grad": [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _p],
    "freud_grad_sumsq": [C.POINTER(TensorList), _p, _p],
    "freud_clip_grads": [C.POINTER(TensorList), _p, _f, _p, _p],
    "freud_adam_step": [C.POINTER(TensorList), _d, _d, _d, _d, _i64, _p, _f, _p],
    "freud_radam_step": [C.POINTER(TensorList), _d, _d, _d, _d, _d, _i64, _p, _f, _p],
    "freud_dp_reduce_scatter": [_p, _p, _i64, _i64, _i64, _i64, _p, _p, _p, _p],
    "freud_dp_adam_allgather": [_p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _p, _p, _p, _i64, _d, _d, _d, _d,
                                _i64, _p, _f, _i, _p],
    "freud_feature_absmax": [_p, _p, _i, _i64, _p, _i64, _p],
    "freud_col_absmax": [_p, _i64, _i64, _p, _p],
    "freud_search_dense": [_p, _i, _p, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _p],
    "freud_search_indexed": [_p, _p, _i, _p, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _p],
    "freud_search_table_dense": [_p, _i, _p, _i64, _i64, _i64, _p, _p, _p, _p],
    "freud_search_topn": [_p, _p, _i64, _i, _i, _d, _i, _d, _i64, _p, _p, _p],
}

# CUDA kernels each entry point enqueues (memsets not counted); bench.py sums these into `gpu_launches`
KERNELS_PER_CALL = {
    "freud_topk_prep_x": 1, "freud_split_operand": 1, "freud_topk_encode_workspace": 0, "freud_topk_encode": 2, "freud_topk_encode_stats": 0, "freud_gemm_nt": 1,
    "freud_row_topk": 1, "freud_gather_rows": 1, "freud_index_map": 1, "freud_row_topk_mask": 1, "freud_col_sum_bf16": 1, "freud_scatter_add_rows": 1, "freud_sum_splits": 1, "freud_gemm_nt_mask": 1, "freud_l1_encode_fused": 1, "freud_l1_decode_fused": 1, "freud_gemm_tn_splitk": 1, "freud_gemm_nn": 1,
    "freud_topk_decode": 1, "freud_topk_dacts": 1, "freud_topk_decode_dacts": 1, "freud_topk_decode_dacts_supported": 0, "freud_topk_refine": 1, "freud_axpby": 1, "freud_shard_merge": 1, "freud_shard_localize": 1, "freud_residual": 1, "freud_csc_build": 5,
    "freud_csc_meta": 1, "freud_topk_sparse_grads": 3, "freud_topk_bdec_grad": 1, "freud_topk_loss_scalars": 1,
    "freud_dead_latent_update": 1, "freud_rownorm_project": 1, "freud_remove_parallel_grad": 1,
    "freud_l1_colnorm": 1, "freud_l1_loss_scalars": 1, "freud_l1_loss_reduce": 1, "freud_l1_dz": 1, "freud_l1_weight_grad": 1, 
    "freud_grad_sumsq": 1, "freud_clip_grads": 1, "freud_adam_step": 1, "freud_radam_step": 1,
    "freud_dp_reduce_scatter": 1, "freud_dp_adam_allgather": 1,
    "freud_feature_absmax": 1, "freud_col_absmax": 1,
    "freud_search_dense": 1, "freud_search_indexed": 1, "freud_search_table_dense": 1, "freud_search_topn": 1,
}

_lib = None
call_count = 0      # C-ABI calls made
kernel_launches = 0  # CUDA kernels those calls enqueued
# when set to a dict, every call is bracketed by CUDA events on the current stream: name -> [(start, end), ...]
profile = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m freud_b200.build` (there is no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        handle.freud_last_error.restype = C.c_char_p
        handle.freud_last_error.argtypes = []
        handle.freud_version.restype = C.c_int
        handle.freud_version.argtypes = []
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        _lib = handle
    return _lib


_fns = {}


def call(name, *args):
    """Invoke a C-ABI entry point; raise RuntimeError(freud_last_error()) on a non-zero status."""
    global call_count, kernel_launches
    fn = _fns.get(name)
    if fn is None:
        fn = _fns[name] = getattr(lib(), name)
    if profile is not None:
        import torch

        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        rc = fn(*args)
        end.record()
        profile.setdefault(name, []).append((start, end))
    else:
        rc = fn(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib().freud_last_error().decode()}")
    call_count += 1
    kernel_launches += KERNELS_PER_CALL[name]


def profile_summary():
    """{name: (calls, total_ms)} from the recorded events (call after torch.cuda.synchronize())."""
    out = {}
    for name, evs in (profile or {}).items():
        out[name] = (len(evs), sum(s.elapsed_time(e) for s, e in evs))
    return out
