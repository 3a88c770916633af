"""ctypes binding of libpixelrec_b200.so (include/pixelrec_b200.h) -- the only way the Python host side
reaches the CUDA kernels.  There is NO fallback: if the library is missing or a call fails, this raises.
"""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpixelrec_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "pixelrec_b200.h")

_P = C.c_void_p
_I64 = C.c_int64
_I = C.c_int
_F = C.c_float
_U64 = C.c_uint64
_U32 = C.c_uint32

# name -> (restype, argtypes); mirrors include/pixelrec_b200.h one to one
SIGNATURES = {
    "pr_version": (_I, []),
    "pr_last_error_string": (C.c_char_p, []),
    "pr_sm_count": (_I, []),
    "pr_set_device": (_I, [_I]),
    "pr_set_tuning": (_I, [_I]),
    "pr_set_seed_device": (_I, [_P]),
    "pr_gather_rows_f32": (_I, [_P, _I64, _I64, _P, _I64, _P, _P, _I, _P]),
    "pr_scatter_plan_workspace_bytes": (C.c_size_t, [_I64, _I64]),
    "pr_scatter_plan": (_I, [_P, _I64, _I64, _I64, _P, _P, _P, _P, _P, _P, C.c_size_t, _P, _P]),
    "pr_scatter_add_rows_f32": (_I, [_P, _I64, _I64, _P, _P, _P, _P, _I64, _F, _P, _P, _P]),
    "pr_adamw_rows_f32": (_I, [_P, _P, _P, _I64, _I64, _P, _P, _F, _F, _F, _F, _F, _F, _I64, _P, _P]),
    "pr_adamw_dense_f32": (_I, [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _F, _I64, _P, _P]),
    "pr_add_ln_fwd_f32": (_I, [_P, _I64, _I64, _P, _I64, _P, _P, _F, _I64, _I64, _F, _F, _U64, _U32, _U32, _P, _P, _P, _P]),
    "pr_add_ln_bwd_partials": (_I, [_I64, _I64]),
    "pr_add_ln_bwd_f32": (_I, [_P, _P, _I64, _I64, _P, _I64, _P, _P, _P, _I64, _I64, _F, _F, _U64, _U32, _U32, _P, _I64, _I,
                               _P, _P, _I, _P]),
    "pr_add_ln_bwd_bias_f32": (_I, [_P, _P, _I64, _I64, _P, _I64, _P, _P, _P, _I64, _I64, _F, _F, _U64, _U32, _U32, _P, _I64, _I,
                                    _P, _P, _I, _P]),
    "pr_add_ln_bwd_bias_z_f32": (_I, [_P, _P, _P, _P, _P, _I64, _I64, _F, _U64, _U32, _P, _P, _P, _I, _P]),
    "pr_act_bwd_bias_partials": (_I, [_I64, _I64]),
    "pr_act_bwd_bias_f32": (_I, [_P, _P, _I64, _I64, _I, _P, _P, _I, _P]),
    "pr_colsum_f32": (_I, [_P, _I, _I, _I64, _P, _P]),
    "pr_colsum_rows_partials": (_I, [_I64, _I64]),
    "pr_colsum_rows_f32": (_I, [_P, _I64, _I64, _P, _I, _P, _P]),
    "pr_act_fwd_f32": (_I, [_P, _I64, _I, _P, _P]),
    "pr_act_bwd_f32": (_I, [_P, _P, _I64, _I, _P, _P]),
    "pr_sasrec_attn_fwd_f32": (_I, [_P, _P, _P, _I64, _P, _I, _I, _I, _I, _I, _F, _U64, _U32, _P, _P, _P]),
    "pr_sasrec_attn_bwd_f32": (_I, [_P, _P, _P, _I64, _P, _P, _I, _I, _I, _I, _I, _F, _U64, _U32, _P, _P, _P, _I64, _P]),
    "pr_sasrec_attn_fwd_tf32": (_I, [_P, _P, _P, _I64, _P, _I, _I, _I, _I, _I, _F, _U64, _U32, _P, _P, _P]),
    "pr_sasrec_attn_bwd_tf32": (_I, [_P, _P, _P, _I64, _P, _P, _I, _I, _I, _I, _I, _F, _U64, _U32, _P, _P, _P, _I64, _P]),
    "pr_bpr_loss_fwd_f32": (_I, [_P, _P, _P, _I64, _P, _I64, _I64, _I64, _P, _P, _P, _P, _P, _P]),
    "pr_bpr_loss_bwd_f32": (_I, [_P, _P, _P, _I64, _P, _P, _I64, _I64, _I64, _P, _P, _P, _I64, _P]),
    "pr_ce_grad_chunk_f32": (_I, [_P, _I64, _I64, _I64, _I64, _P, _P, _P, _I, _P]),
    "pr_seq_batch_build": (_I, [_P, _I64, _I, _P, _I64, _I64, _U64, _P, _P, _P, _P]),
    "pr_score_topk_workspace_bytes": (C.c_size_t, [_I64, _I64, _I]),
    "pr_score_topk_f32": (_I, [_P, _I64, _P, _I64, _I64, _P, _P, _I64, _I, _I, _P, _P, _P, C.c_size_t, _P]),
    "pr_score_topk_exact_workspace_bytes": (C.c_size_t, [_I64, _I64, _I]),
    "pr_table_norm_max_f32": (_I, [_P, _I64, _I64, _P, _P]),
    "pr_score_topk_exact_f32": (_I, [_P, _I64, _P, _I64, _I64, _P, _P, _I64, _I, _I, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "pr_gemm_colsum_rows": (_I, [_I64]),
    "pr_gemm_tf32": (_I, [_P, _I, _I64, _P, _I, _I64, _I64, _I64, _I64, _P, _P, _I, _I, _P, _P, _I, _P, _I, _P]),
    "pr_gemm_tf32_drop": (_I, [_P, _I, _I64, _P, _I, _I64, _I64, _I64, _I64, _P, _P, _P, _I, _F, _U64, _U32, _P]),
    "pr_gemm_splitk_reduce_f32": (_I, [_P, _I, _I64, _P, _P]),
    "pr_score_prepare_f16": (_I, [_P, _I64, _P, _P, _P]),
    "pr_score_topk_f16_workspace_bytes": (C.c_size_t, [_I64, _I64, _I64, _I]),
    "pr_score_topk_f16": (_I, [_P, _I64, _P, _I64, _I64, _P, _P, _I64, _I, _I, _P, _P, _P, C.c_size_t, _P, _P]),
    "pr_score_ce_workspace_bytes": (C.c_size_t, [_I64, _I64]),
    "pr_score_ce_f32": (_I, [_P, _I64, _P, _I64, _I64, _P, _I, _P, _P, _P, _P, C.c_size_t, _P]),
    "pr_attn_long_fwd_f32": (_I, [_P, _P, _P, _I64, _P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "pr_attn_long_bwd_f32": (_I, [_P, _P, _P, _I64, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _I64, _P, _P]),
    "pr_shared_alloc": (_I, [C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]),
    "pr_shared_free": (_I, [_P]),
    "pr_shared_open": (_I, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "pr_shared_close": (_I, [_P]),
    "pr_gather_rows_peers_f32": (_I, [_P, _I, _I64, _I64, _P, _I64, _P, _P, _P]),
    "pr_peer_barrier": (_I, [_P, _I, _I, _U64, _P, _P, _P]),
    "pr_plan_inverse": (_I, [_P, _P, _P, _I64, _I64, _P, _P]),
    "pr_gather_rows_peers_plan_f32": (_I, [_P, _I, _I64, _I64, _P, _P, _I64, _I64, _I64, _P, _P, _P]),
    "pr_push_rows_peers_plan_f32": (_I, [_P, _P, _P, _I64, _I64, _I, _I, _I64, _P, _P, _P, _P, _P]),
    "pr_push_rows_peers_f32": (_I, [_P, _P, _I64, _I64, _I, _I, _I64, _I64, _P, _P, _P, _P, _P]),
}


class PixelRecB200Error(RuntimeError):
    pass


def header_symbols():
    """Every entry point include/pixelrec_b200.h declares (used by the CPU-side ABI test)."""
    with open(HEADER_PATH) as f:
        return re.findall(r"PR_API\s+[\w\s\*]+?\b(pr_[a-z0-9_]+)\s*\(", f.read())


_lib = None


def load():
    """Loads the shared library (no GPU needed to load).  Raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PixelRecB200Error(
            f"{LIB_PATH} not found: the CUDA extension is mandatory (no CPU / eager fallback). "
            "Build it with `python -m pixelrec_b200.build` or `__graft_entry__.build()`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.pr_version() != 1:
        raise PixelRecB200Error(f"ABI version mismatch: library reports {lib.pr_version()}")
    _lib = lib
    return lib


def check(rc, name="call"):
    if rc != 0:
        msg = load().pr_last_error_string().decode("utf-8", "replace")
        kind = "invalid argument" if rc == -1 else ("unsupported" if rc == -2 else f"cudaError {rc}")
        raise PixelRecB200Error(f"{name} failed ({kind}): {msg}")
