"""ctypes binding of include/mst_b200.h (libmst_b200.so).

This is the stub a maintainer of the reference would add next to `mst/models/dino.py` (see
INTEGRATION.md).  The library is loaded from this directory; if it has not been built the import
fails loudly -- there is no Python/CPU fallback for any compute entry point.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MST_LIB_PATH") or os.path.join(_HERE, "libmst_b200.so")   # MST_LIB_PATH: A/B builds
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "mst_b200.h")

PRECISION = {"fp32": 0, "bf16": 1}
FUSION = {"transformer": 0, "linear": 1, "average": 2}
SRC_DTYPE = {"torch.float32": 0, "torch.bfloat16": 1, "torch.float16": 2}
ABI_VERSION = 4


class MSTError(RuntimeError):
    pass


class MstConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("embed_dim", "depth", "enc_heads", "slice_heads", "out_ch", "pos_tokens", "precision", "device",
                 "num_registers", "use_bottleneck", "use_slice_pos_emb", "slice_fusion", "enable_linear", "rotary",
                 "interpolate_antialias")] + [("interpolate_offset", ctypes.c_float)]


def declared_symbols():
    """Function names declared in include/mst_b200.h."""
    with open(HEADER_PATH) as f:
        return re.findall(r"MST_API\s+[\w\s\*]+?\b(mst_\w+)\s*\(", f.read())


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python new-vit_b200/build.py` "
            "(or __graft_entry__.build()). There is no fallback path.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t
    L.mst_abi_version.restype = ctypes.c_int
    L.mst_last_error.restype = ctypes.c_char_p
    L.mst_create.argtypes = [ctypes.POINTER(MstConfig), ctypes.POINTER(vp)]
    L.mst_destroy.argtypes = [vp]
    L.mst_set_weight.argtypes = [vp, ctypes.c_char_p, vp, i64, vp]
    L.mst_set_weights.argtypes = [vp, i32, vp, vp, vp, vp]
    L.mst_finalize_weights.argtypes = [vp, vp]
    L.mst_workspace_bytes.argtypes = [vp, i32, i32, i32, i32, ctypes.POINTER(sz)]
    L.mst_forward.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    L.mst_slice_head_forward.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]
    L.mst_saliency.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    L.mst_rollout.argtypes = [vp, i32, i32, i32, vp, vp, vp]
    L.mst_pos_embed.argtypes = [vp, i32, i32, vp, vp]
    L.mst_quantile_workspace_bytes.argtypes = [i32, i32, ctypes.POINTER(sz)]
    L.mst_quantile.argtypes = [vp, i64, i32, vp, i32, vp, vp, sz, vp]
    L.mst_prepare_volume_workspace_bytes.argtypes = [i32, i32, i32, i32, ctypes.POINTER(sz)]
    L.mst_prepare_volume.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, ctypes.c_float, ctypes.c_float, vp, vp, vp, sz, vp]
    L.mst_kernel_gemm_bf16.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]
    L.mst_kernel_gemm_bf16_ln.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]
    L.mst_kernel_pack_linear_ln.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, vp]
    L.mst_kernel_gemm_bf16_res_stats.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, ctypes.c_float, vp]
    L.mst_kernel_gemm_bf16_f32out.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    L.mst_kernel_wgrad_bf16.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
    L.mst_kernel_ln_bwd_bf16.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, ctypes.c_float, vp]
    L.mst_kernel_gelu_bf16.argtypes = [vp, vp, vp, vp, i64, vp]
    L.mst_kernel_transpose_bf16.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    L.mst_kernel_attention_bwd_bf16.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp]
    L.mst_kernel_attention_lse_bf16.argtypes = [vp, vp, vp, i32, i32, vp]
    L.mst_kernel_attention_bwd_lse_bf16.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp]
    L.mst_kernel_row_stats_bf16.argtypes = [vp, vp, i32, i32, ctypes.c_float, vp]
    L.mst_debug_gemm_timing.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp]
    L.mst_kernel_gemm_f32.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]
    L.mst_kernel_attention_bf16.argtypes = [vp, vp, i32, i32, i32, vp]
    L.mst_kernel_attention_bf16_warp_mma.argtypes = [vp, vp, i32, i32, i32, vp]
    L.mst_kernel_attention_f32.argtypes = [vp, vp, i32, i32, i32, vp]
    L.mst_kernel_layernorm_bf16.argtypes = [vp, vp, vp, vp, i32, i32, ctypes.c_float, vp]
    fl = ctypes.c_float
    L.mst_slice_train_bytes.argtypes = [i32, i32, i32, i32, i32, ctypes.POINTER(sz), ctypes.POINTER(sz)]
    L.mst_slice_train_forward.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]
    L.mst_slice_train_backward.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]
    L.mst_adamw.argtypes = [vp, vp, vp, vp, vp, i64, fl, fl, fl, fl, fl, i32, fl, vp]
    L.mst_train_workspace_bytes.argtypes = [vp, i32, i32, i32, i32, ctypes.POINTER(sz)]
    L.mst_train_forward.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp, sz, vp]
    L.mst_set_grad.argtypes = [vp, ctypes.c_char_p, vp, i64]
    L.mst_train_backward.argtypes = [vp, vp, i32, i32, i32, i32, vp, sz, vp]
    L.mst_profile_begin.argtypes = [vp]
    L.mst_profile_end.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64), i32]
    L.mst_launch_count.argtypes = [vp]
    L.mst_set_graph_threshold.argtypes = [vp, i64]
    L.mst_graph_replays.argtypes = [vp]
    for name in declared_symbols():
        fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
        if name not in ("mst_last_error", "mst_profile_categories", "mst_launch_count", "mst_graph_replays"):
            fn.restype = ctypes.c_int
    L.mst_profile_categories.restype = ctypes.c_char_p
    L.mst_launch_count.restype = ctypes.c_ulonglong
    L.mst_graph_replays.restype = ctypes.c_ulonglong
    if L.mst_abi_version() != ABI_VERSION:
        raise ImportError("libmst_b200.so ABI version mismatch")
    _lib = L
    return L


def check(status):
    if status != 0:
        raise MSTError(lib().mst_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())
