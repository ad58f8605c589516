"""ctypes binding of libisp_b200.so (the C ABI declared in include/isp_b200.h).

There is no CPU fallback and no alternative backend: if the shared library is
missing, or a tensor is not on a CUDA device, the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ISP_B200_LIB") or os.path.join(_HERE, "libisp_b200.so")

_c = ctypes
_P, _I, _F, _LL, _S = _c.c_void_p, _c.c_int, _c.c_float, _c.c_longlong, _c.c_void_p

# name -> argtypes; must list every symbol include/isp_b200.h declares (tests check this)
SIGNATURES = {
    "isp_distmaps_fwd": [_P, _P, _I, _I, _I, _I, _F, _F, _I, _S],
    "isp_distmaps_rounded_sqdist_fwd": [_P, _P, _I, _I, _I, _I, _F, _S],
    "isp_prepare_input_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _F, _F, _I, _S],
    "isp_nchw_to_nhwc_f32": [_P, _P, _I, _I, _I, _I, _LL, _LL, _LL, _LL, _S],
    "isp_nchw_to_nhwc_bf16": [_P, _P, _I, _I, _I, _I, _I, _LL, _LL, _LL, _LL, _S],
    "isp_bilinear_ac_nhwc": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _S],
    "isp_bilinear_ac_nhwc_bwd": [_P, _P, _I, _I, _I, _I, _I, _I, _S],
    "isp_bilinear_ac_nhwc_dual": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _S],
    "isp_bilinear_ac_nhwc_bias": [_P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _S],
    "isp_jbu_pool_guidance": [_P, _P, _I, _I, _I, _I, _I, _LL, _LL, _LL, _LL, _S],
    "isp_jbu_range_proj": [_P, _P, _LL, _P, _P, _P, _P, _S],
    "isp_jbu_filters": [_P, _P, _P, _I, _I, _I, _F, _F, _P, _P, _P, _P, _I, _S],
    "isp_jbu_filters_simt": [_P, _P, _P, _I, _I, _I, _F, _F, _P, _P, _P, _P, _I, _S],
    "isp_jbu_bicubic2x_reflectpad": [_P, _P, _I, _I, _I, _I, _S],
    "isp_jbu_bicubic2x_reflectpad_bwd": [_P, _P, _I, _I, _I, _I, _S],
    "isp_adaptive_conv_fwd": [_P, _P, _P, _I, _I, _I, _I, _I, _S],
    "isp_adaptive_conv_grad_input": [_P, _P, _P, _I, _I, _I, _I, _S],
    "isp_gemm_f32_simt": [_P, _P, _P, _P, _F, _P, _LL, _I, _I, _S],
    "isp_gemm_bf16_tc": [_P, _LL, _P, _LL, _P, _P, _I, _LL, _F, _I, _P, _LL, _I, _LL, _I, _I, _S],
    "isp_conv3x3_bf16_tc": [_P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _S],
    "isp_gemm_stats_slots": [_I, _I, _I],
    "isp_gemm_bf16_tc_ex": [_P, _LL, _P, _LL, _P, _P, _I, _LL, _F, _I, _P, _LL, _I, _LL, _I, _I, _P, _I, _P, _F, _P, _I, _S],
    "isp_conv3x3_bf16_tc_ex": [_P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _S],
    "isp_conv3x3_dgrad_bf16_tc": [_P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _S],
    "isp_conv3x3_wgrad_bf16_tc": [_P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _S],
    "isp_head_classifier_bwd": [_P, _LL, _P, _P, _P, _LL, _P, _P, _P, _LL, _I, _S],
    "isp_colsum_bf16": [_P, _LL, _P, _LL, _I, _S],
    "isp_ffn_fused_bf16_tc": [_P, _LL, _I, _P, _LL, _P, _P, _I, _P, _LL, _P, _I, _P, _LL, _LL, _P, _I, _F, _P, _S],
    "isp_zoom_in_fwd": [_P, _P, _I, _I, _I, _I, _I, _I, _P, _I, _I, _I, _S],
    "isp_unzoom_probs": [_P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _F, _F, _P, _P, _S],
    "isp_noc_next_click": [_P, _P, _P, _I, _I, _P, _P, _S],
    "isp_col_moments_bf16": [_P, _LL, _LL, _I, _P, _I, _S],
    "isp_bn_relu_rows_bf16": [_P, _LL, _P, _P, _LL, _I, _P, _S],
    "isp_gemm_bf16_tc_batched": [_P, _LL, _LL, _LL, _P, _LL, _LL, _LL, _P, _LL, _LL, _LL, _I, _I, _I, _I, _I, _I, _F, _S],
    "isp_gemm_bf16_tc_batched_tn": [_P, _LL, _LL, _LL, _P, _LL, _LL, _LL, _P, _LL, _LL, _LL, _I, _I, _I, _I, _I, _I, _F, _S],
    "isp_gemm_bf16_tc_batched_nn": [_P, _LL, _LL, _LL, _P, _LL, _LL, _LL, _P, _LL, _LL, _LL, _I, _I, _I, _I, _I, _I, _F, _S],
    "isp_layernorm_rows_bwd": [_P, _LL, _P, _I, _LL, _P, _P, _LL, _P, _LL, _P, _LL, _LL, _I, _F, _S],
    "isp_layernorm_affine_bwd": [_P, _LL, _P, _I, _LL, _P, _P, _LL, _I, _F, _S],
    "isp_gelu_bwd_bf16": [_P, _P, _P, _LL, _I, _S],
    "isp_softmax_rows": [_P, _LL, _P, _LL, _LL, _I, _I, _S],
    "isp_attn_ds_rows": [_P, _LL, _P, _I, _LL, _P, _LL, _LL, _I, _I, _S],
    "isp_transpose_bf16_batched": [_P, _LL, _LL, _P, _LL, _LL, _I, _I, _I, _S],
    "isp_attention_bf16_tc": [_P, _LL, _I, _P, _P, _P, _LL, _I, _I, _LL, _I, _I, _I, _S],
    "isp_attention_bf16_tc_lse": [_P, _LL, _I, _P, _P, _P, _LL, _I, _I, _LL, _I, _I, _I, _P, _S],
    "isp_attention_bf16_tc_opt": [_P, _LL, _I, _P, _P, _P, _LL, _I, _I, _LL, _I, _I, _I, _P, _I, _I, _S],
    "isp_attention_rowdot_heads": [_P, _LL, _P, _LL, _P, _I, _LL, _I, _I, _S],
    "isp_attention_bwd_bf16_tc": [_P, _LL, _P, _LL, _P, _P, _P, _P, _P, _P, _P, _LL, _I, _LL, _I, _I, _I, _S],
    "isp_layernorm_rows": [_P, _I, _LL, _P, _I, _LL, _P, _P, _LL, _I, _F, _S],
    "isp_minmax_per_channel": [_P, _P, _I, _I, _I, _LL, _LL, _S],
    "isp_loftup_fourier_chnorm": [_P, _LL, _LL, _LL, _LL, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _S],
    "isp_loftup_lr_prepare": [_P, _LL, _LL, _LL, _LL, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _S],
    "isp_repack_heads": [_P, _I, _LL, _I, _I, _P, _I, _I, _I, _I, _I, _I, _S],
    "isp_conv3x3_s2_c32": [_P, _LL, _LL, _LL, _LL, _P, _P, _P, _I, _I, _I, _I, _S],
    "isp_conv3x3_s2_c32_raw": [_P, _LL, _LL, _LL, _LL, _P, _P, _P, _I, _I, _I, _I, _S],
    "isp_adaptive_maxpool_nhwc": [_P, _P, _I, _I, _I, _I, _I, _I, _S],
    "isp_copy_channels": [_P, _I, _LL, _LL, _LL, _LL, _P, _I, _LL, _LL, _LL, _I, _I, _I, _I, _S],
    "isp_rowdot": [_P, _I, _LL, _P, _P, _P, _LL, _I, _I, _S],
    "isp_vit_patchify": [_P, _LL, _LL, _LL, _LL, _P, _I, _I, _I, _I, _I, _I, _S],
    "isp_vit_assemble_tokens": [_P, _P, _P, _P, _P, _I, _I, _I, _S],
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C isegprobe_b200/csrc). isegprobe_b200 has no CPU or PyTorch fallback.")
        L = ctypes.CDLL(LIB_PATH)
        L.isp_version.restype = _I
        L.isp_last_error.restype = _c.c_char_p
        L.isp_launch_count.restype = _c.c_ulonglong
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _I
        _lib = L
    return _lib


def launch_count() -> int:
    return int(lib().isp_launch_count())


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def dptr(t):
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("isegprobe_b200 kernels need CUDA tensors (there is no CPU fallback)")
    return t.data_ptr()


def call(name, *args):
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib().isp_last_error().decode()}")
