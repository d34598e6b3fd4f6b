"""ctypes binding of libltx_b200.so (C ABI in include/ltx_b200.h).

No fallback: if the shared library is missing or a call fails, this raises.  Build it with
`python -c "import __graft_entry__ as g; g.build()"` or `ltx-video-gpupoor_b200/csrc/build.sh`.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_longlong, c_size_t, c_uint, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# LTXB200_LIB selects another build of the same ABI (e.g. the -DLTXB200_DEBUG_HANG watchdog build used while developing)
LIB_PATH = os.environ.get("LTXB200_LIB") or os.path.join(_HERE, "libltx_b200.so")

_lib = None


class LtxB200Error(RuntimeError):
    pass


def _declare(lib):
    P, I, L, F = c_void_p, c_int, c_int64, c_float
    sig = {
        "ltxb200_abi_version": ([], c_int),
        "ltxb200_error_string": ([I], c_char_p),
        "ltxb200_launch_count": ([], c_longlong),
        "ltxb200_gemm_bf16": ([P, L, P, L, I, I, I, P, L, I, P, I, P, L, P, L, I, P], I),
        "ltxb200_conv3d_bf16": ([P, P, P, P, I, I, I, I, I, I, I, I, I, P, P], I),
        "ltxb200_attention_bf16": ([P, L, L, P, L, L, P, L, L, P, L, L, I, I, I, I, I, F, P, P], I),
        "ltxb200_attention_klens_bf16": ([P, L, L, P, L, L, P, L, L, P, L, L, I, I, I, I, I, F, P, P], I),
        "ltxb200_attention_acc_bf16": ([P, L, L, P, L, L, P, L, L, P, L, L, I, I, I, I, I, F, P, P], I),
        "ltxb200_norm_mod_bf16": ([P, L, P, L, I, I, P, P, L, I, P, P, F, I, P], I),
        "ltxb200_qk_norm_rope_bf16": ([P, L, I, P, L, I, I, P, P, P, P, I, F, P], I),
        "ltxb200_qk_norm_rope_wan_bf16": ([P, L, I, P, L, I, I, P, P, P, P, I, I, I, F, P], I),
        "ltxb200_lincomb_f32": ([P, L, I, P, P, P], I),
        "ltxb200_rf_step_tokens_f32": ([P, P, P, P, P, L, I, P, I, P], I),
        "ltxb200_ada_add_bf16": ([P, P, P, I, I, I, P], I),
        "ltxb200_act_bf16": ([P, P, L, I, P], I),
        "ltxb200_stg_blend_bf16": ([P, P, L, P, I, L, I, P], I),
        "ltxb200_axpby_bf16": ([P, P, P, L, F, F, P], I),
        "ltxb200_rel_l1_bf16": ([P, P, L, P, P], I),
        "ltxb200_groupnorm_silu_bf16": ([P, P, I, L, I, P, P, P, F, I, P, P], I),
        "ltxb200_adain_f32": ([P, P, P, I, L, L, F, P], I),
        "ltxb200_latent_from_ndhwc": ([P, P, I, I, L, P, P, P], I),
        "ltxb200_bilinear_resize_f32": ([P, P, L, I, I, I, I, P], I),
        "ltxb200_timestep_embed": ([P, P, I, I, P], I),
        "ltxb200_cast_f32_to_bf16": ([P, P, L, P], I),
        "ltxb200_guidance_step": ([P, L, L, I, I, I, I, F, F, F, P, P, P, I, F, P, P, P], I),
        "ltxb200_guidance_step_stochastic": ([P, L, L, I, I, I, I, F, F, F, P, P, P, I, F, P, P, P, P], I),
        "ltxb200_cfg_combine_f32": ([P, P, P, L, F, I, P, P], I),
        "ltxb200_pixelnorm_silu_bf16": ([P, P, L, I, F, I, P], I),
        "ltxb200_pixelnorm_mod_silu_bf16": ([P, P, L, I, F, P, P, I, P], I),
        "ltxb200_latent_to_ndhwc": ([P, I, P, I, I, L, P, P, P], I),
        "ltxb200_gemm_bf16_f32res": ([P, L, P, L, I, I, I, P, L, P, I, P, L, P, L, I, P], I),
        "ltxb200_norm_mod_f32in": ([P, L, P, L, I, I, P, P, L, I, F, I, P], I),
        "ltxb200_ada_add_f32": ([P, P, P, I, I, I, P], I),
        "ltxb200_conv3d_norm_bf16": ([P, P, P, P, P, I, I, I, I, I, I, I, P, I, F, P], I),
        "ltxb200_conv3d_strided_bf16": ([P, P, P, P, I, I, I, I, I, I, I, I, P], I),
        "ltxb200_conv_taps_bf16": ([P, P, P, P, I, I, I, I, I, I, I, I, I, P, P], I),
        "ltxb200_conv_taps_strided_bf16": ([P, P, P, P, I, I, I, I, I, I, I, I, I, I, I, P], I),
        "ltxb200_l2norm_silu_bf16": ([P, P, L, I, I, P, I, P], I),
        "ltxb200_upsample2x_nhwc_bf16": ([P, P, L, I, I, I, P], I),
        "ltxb200_softmax_rows_f32_bf16": ([P, L, P, L, I, I, F, P], I),
        "ltxb200_comm_alloc": ([c_size_t, POINTER(c_void_p), P], I),
        "ltxb200_comm_open": ([P, POINTER(c_void_p)], I),
        "ltxb200_comm_close": ([P], I),
        "ltxb200_comm_free": ([P], I),
        "ltxb200_comm_wait": ([P, I, c_uint, P], I),
        "ltxb200_comm_wait_status": ([P, I, c_uint, P, P, c_uint, P], I),
        "ltxb200_peer_allgather": ([P, L, I, I, I, POINTER(c_void_p), POINTER(c_void_p), c_uint, P, P], I),
        "ltxb200_qk_norm_rope_wan_scatter_rows_bf16": ([P, L, I, I, I, I, P, P, P, P, I, I, I, F, I, I, I, POINTER(c_void_p), POINTER(c_void_p), c_uint, P, c_uint, I, P], I),
        "ltxb200_gemm_qkv_vscatter_bf16": ([P, L, P, L, I, I, I, P, L, P, I, I, I, I, I, I, I, POINTER(c_void_p), P], I),
        "ltxb200_scatter_signal_ctas": ([I], c_uint),
        "ltxb200_qk_norm_rope_wan_scatter_bf16": ([P, L, I, I, P, P, P, P, I, I, I, F, I, I, I, POINTER(c_void_p), POINTER(c_void_p), c_uint, P, P], I),
        "ltxb200_attention_scatter_bf16": ([P, L, L, P, L, L, P, L, L, L, I, I, I, I, I, F, P, I, I, POINTER(c_void_p), POINTER(c_void_p), c_uint, P, I, I, P], I),
    }
    for name, (args, res) in list(sig.items()):
        if os.environ.get("LTXB200_LIB") and not hasattr(lib, name):      # older development build selected by hand
            del sig[name]
            continue
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    return sig


EXPORTED_SYMBOLS = None


def lib():
    global _lib, EXPORTED_SYMBOLS
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LtxB200Error(
                f"{LIB_PATH} not found: the sm_100a CUDA library is not built. There is no CPU/PyTorch fallback; "
                "run ltx-video-gpupoor_b200/csrc/build.sh (or __graft_entry__.build()).")
        l = ctypes.CDLL(LIB_PATH)
        EXPORTED_SYMBOLS = sorted(_declare(l).keys())
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().ltxb200_error_string(rc).decode()
        raise LtxB200Error(f"libltx_b200 {what} failed: {msg} (code {rc})")


def launch_count() -> int:
    return int(lib().ltxb200_launch_count())
