"""ctypes binding of libtvae_b200.so (the C ABI declared in include/tvae.h).

The product path has no fallback: if the shared library is missing this module raises at import time, and every
wrapper raises on a non-zero return code with the library's own error message.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtvae_b200.so")


class TvaeError(RuntimeError):
    pass


class ConvArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("x_pitch", C.c_int32),
        ("kind", C.c_int32), ("R", C.c_int32), ("flip", C.c_int32),
        ("w", C.c_void_p),
        ("w_rows", C.c_int32), ("k_pitch", C.c_int32), ("c_pad", C.c_int32),
        ("Cout", C.c_int32),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p),
        ("res_pitch", C.c_int32),
        ("out_f32", C.c_void_p),
        ("out_f32_pitch", C.c_int32),
        ("out_bf16", C.c_void_p),
        ("out_bf16_pitch", C.c_int32),
        ("bn", C.c_int32),
        ("stats_part", C.c_void_p),
        ("stats_groups", C.c_int32),
        ("x_lo", C.c_void_p),
        ("w_lo", C.c_void_p),
        ("out_bf16_lo", C.c_void_p),
        ("nll_x", C.c_void_p),
        ("nll_x_pitch", C.c_int32),
        ("nll_loss_type", C.c_int32),
        ("nll_logvar", C.c_void_p),
        ("nll_batch", C.c_int32),
        ("nll_workspace", C.c_void_p),
        ("nll_sums", C.c_void_p),
    ]


class WgradArgs(C.Structure):
    _fields_ = [
        ("p", C.c_void_p),
        ("p_pitch", C.c_int32), ("Cm", C.c_int32),
        ("q", C.c_void_p),
        ("q_pitch", C.c_int32), ("Cn", C.c_int32),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("kind", C.c_int32), ("R", C.c_int32),
        ("splits", C.c_int32),
        ("workspace", C.c_void_p),
        ("grad", C.c_void_p),
        ("accumulate", C.c_int32),
        ("flip", C.c_int32),
        ("grad_ld", C.c_int32), ("grad_off", C.c_int32),
    ]


def _load():
    if not os.path.exists(LIB_PATH):
        raise TvaeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C tempo_vae_b200/csrc`). There is no CPU/PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float
    u32 = C.c_uint32
    sigs = {
        "tvae_last_error": (C.c_char_p, []),
        "tvae_abi_version": (i32, []),
        "tvae_conv_gemm": (i32, [C.POINTER(ConvArgs), vp]),
        "tvae_conv_nll_workspace_bytes": (i64, [i64, i32]),
        "tvae_conv_set_cta_pair": (i32, [i32]),
        "tvae_conv_set_trace": (i32, [vp, i32]),
        "tvae_wgrad_gemm": (i32, [C.POINTER(WgradArgs), vp]),
        "tvae_wgrad_set_cta_pair": (i32, [i32]),
        "tvae_wgrad_workspace_bytes": (i64, [i32, i32, i32, i32]),
        "tvae_wgrad_splits": (i32, [i32, i32, i32, i64]),
        "tvae_wgrad_skinny_workspace_bytes": (i64, [i32]),
        "tvae_wgrad_skinny": (i32, [vp, i32, i32, vp, i32, i32, i32, i32, i32, i32, vp, i64, i64, i32, vp, vp]),
        "tvae_pack_chunk_elems": (i32, []),
        "tvae_pack_weights_batched": (i32, [vp, vp, i32, i64, vp]),
        "tvae_pack_weight": (i32, [vp, vp, i32, i32, i32, i32, i32, i64, i64, i64, vp, vp]),
        "tvae_nchw_f32_to_nhwc_bf16": (i32, [vp, vp, i32, i32, i32, i32, vp, vp]),
        "tvae_normalize_radiance": (i32, [vp, vp, vp, i64, i32, f32, f32, f32, vp, vp, i32, vp]),
        "tvae_nhwc_f32_to_nhwc_bf16": (i32, [vp, i64, i64, i32, vp, i32, vp, vp]),
        "tvae_nhwc_f32_to_nchw_f32": (i32, [vp, vp, i32, i32, i32, i32, vp]),
        "tvae_nhwc_bf16_to_nchw_f32": (i32, [vp, vp, i32, i32, i32, i32, vp]),
        "tvae_f32_to_bf16": (i32, [vp, vp, i64, vp, vp]),
        "tvae_gn_stats": (i32, [vp, i32, i32, i32, i32, f32, vp, vp]),
        "tvae_gn_stats_finalize": (i32, [vp, i32, i32, i32, C.c_double, f32, vp, vp]),
        "tvae_gn_act_fwd": (i32, [vp, i32, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]),
        "tvae_gn_act_fwd2": (i32, [vp, i32, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp]),
        "tvae_gn_act_bwd2": (i32, [vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]),
        "tvae_gn_bwd_workspace_bytes": (i64, [i32, i32, i32, i32]),
        "tvae_gn_act_bwd": (i32, [vp, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]),
        "tvae_gn_set_bwd_fused": (i32, [i32, i32]),
        "tvae_colsum_workspace_bytes": (i64, [i64, i32]),
        "tvae_colsum_bf16": (i32, [vp, i64, i32, i32, vp, vp, vp]),
        "tvae_attn_fwd": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp]),
        "tvae_attn_bwd": (i32, [vp, vp, vp, i32, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp]),
        "tvae_attn_set_tcgen05": (i32, [i32]),
        "tvae_attn_fwd_tc": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp]),
        "tvae_attn_bwd_tc": (i32, [vp, vp, vp, i32, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp]),
        "tvae_reparam_fwd": (i32, [vp, vp, u64, u64, i32, i32, i32, vp, i32, vp, vp, vp, vp, vp]),
        "tvae_reparam_bwd": (i32, [vp, vp, vp, vp, vp, f32, i32, i32, i32, vp, vp]),
        "tvae_recon_metrics_workspace_bytes": (i64, [i32]),
        "tvae_recon_metrics": (i32, [vp, i32, vp, i32, i32, i32, i32, vp, vp, vp]),
        "tvae_nll_workspace_bytes": (i64, [i32]),
        "tvae_nll_fwd": (i32, [vp, i32, vp, i32, i64, i32, i32, vp, i32, vp, i32, vp, vp, vp, vp]),
        "tvae_vae_loss_finalize": (i32, [vp, vp, i32, vp, C.c_double, f32, vp, vp]),
        "tvae_l2head_loss_fwd": (i32, [vp, i32, C.POINTER(vp), i32, i32, i32, i32, vp, vp]),
        "tvae_l2head_loss_bwd": (i32, [vp, i32, C.POINTER(vp), i32, i32, i32, i32, vp, vp, f32, vp, i32, vp]),
        "tvae_l2head_finalize": (i32, [vp, vp, i32, vp, vp, vp]),
        "tvae_sumsq_workspace_bytes": (i64, [i64]),
        "tvae_sumsq": (i32, [vp, i64, vp, vp, vp]),
        "tvae_adamw": (i32, [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i64, vp, f32, f32, vp]),
        "tvae_gather_rows": (i32, [vp, i64, i64, vp, i32, vp, vp]),
        "tvae_extract_tiles": (i32, [vp, i32, i32, i32, vp, i32, i32, vp, vp, f32, f32, f32, vp, vp, i32, vp]),
        "tvae_spectrum_stats_workspace_bytes": (i64, [i64, i32]),
        "tvae_spectrum_stats_accum": (i32, [vp, i64, i32, f32, i32, vp, vp, vp]),
        "tvae_spectrum_stats_finalize": (i32, [vp, i64, i32, vp, vp, vp]),
        "tvae_batch_stats_workspace_bytes": (i64, []),
        "tvae_batch_stats": (i32, [vp, i32, i64, i64, i64, vp, vp, vp]),
        "tvae_act_dropout_fwd": (i32, [vp, i32, i64, i32, i32, f32, u64, u64, vp, i32, vp]),
        "tvae_act_dropout_bwd": (i32, [vp, i32, vp, i32, i64, i32, i32, f32, u64, u64, vp, i32, vp]),
        "tvae_probe_mse": (i32, [vp, i32, vp, i64, i64, i64, vp, vp, i32, vp]),
        "tvae_nan_moments": (i32, [vp, i64, f32, vp, vp]),
        "tvae_select_hist": (i32, [vp, i64, f32, i32, u32, u32, i32, vp, vp]),
        "tvae_component_pool": (i32, [vp, i32, i32, i32, i32, i32, f32, f32, vp, vp, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib, tuple(sigs)


lib, EXPORTED = _load()
if os.environ.get("TVAE_WGRAD_CTA_PAIR") in ("0", "1"):
    lib.tvae_wgrad_set_cta_pair(int(os.environ["TVAE_WGRAD_CTA_PAIR"]))
if os.environ.get("TVAE_CONV_CTA_PAIR") in ("0", "1"):     # A/B switch of the conv schedule (results are bit-identical)
    lib.tvae_conv_set_cta_pair(int(os.environ["TVAE_CONV_CTA_PAIR"]))
if os.environ.get("TVAE_ATTN_TCGEN05") in ("0", "1"):      # A/B switch: tcgen05 kind::tf32 attention vs the mma.sync kernels
    lib.tvae_attn_set_tcgen05(int(os.environ["TVAE_ATTN_TCGEN05"]))


def last_error() -> str:
    return lib.tvae_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise TvaeError(f"{what} failed (rc={rc}): {last_error()}")
