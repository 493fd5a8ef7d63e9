"""ctypes binding of libsei_b200.so (include/sei_b200.h).  No CPU fallback: if the library or
a CUDA device is missing every operator raises."""
import ctypes as C
import os

from .build import LIB_PATH

_lib = None

_vp, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float
_ip = C.POINTER(C.c_int)

SIGNATURES = {
    "sei_abi_version": (C.c_int, []),
    "sei_last_error": (C.c_char_p, []),
    "sei_last_kernel": (C.c_char_p, []),
    "sei_launch_count": (C.c_longlong, []),
    "sei_device_info": (C.c_int, [_ip, _ip, _ip, _ip]),
    "sei_blur_circular_f32": (C.c_int, [_vp, _vp, _ll, _i, _i, _vp, _i, _i, _i, _vp, _f, _i, _vp]),
    "sei_blur_padded_f32": (C.c_int, [_vp, _vp, _ll, _i, _i, _vp, _i, _i, _i, _i, _ip, _ip, _vp]),
    "sei_down_aa_f32": (C.c_int, [_vp, _vp, _ll, _i, _i, _i, _vp, _f, _i, _vp]),
    "sei_down_aa_transpose_f32": (C.c_int, [_vp, _vp, _ll, _i, _i, _i, _i, _vp]),
    "sei_up_bicubic_f32": (C.c_int, [_vp, _vp, _ll, _i, _i, _i, _vp]),
    "sei_crop_batch_f32": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "sei_scale_transform_f32": (C.c_int, [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp]),
    "sei_scale_transform_src_f32": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "sei_scale_transform_src_backward_f32": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "sei_scale_transform_backward_f32": (C.c_int, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "sei_scale_params_f32": (C.c_int, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "sei_ei_workspace_bytes": (C.c_longlong, [_i, _i]),
    "sei_ei_remeasure_f32": (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _f, _vp, _vp]),
    "sei_reduce_workspace_bytes": (C.c_longlong, []),
    "sei_mse_f32": (C.c_int, [_vp, _vp, _ll, _vp, _vp, _vp]),
    "sei_mse_backward_f32": (C.c_int, [_vp, _vp, _ll, _vp, _vp, _vp, _vp]),
    "sei_sure_loss_f32": (C.c_int, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _f, _i, _vp, _vp, _vp]),
    "sei_sure_loss_backward_f32": (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _f, _vp, _vp, _vp, _vp]),
    "sei_sure_perturb_f32": (C.c_int, [_vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp]),
    "sei_roll_f32": (C.c_int, [_vp, _vp, _ll, _i, _i, _i, _i, _vp]),
    "sei_add_noise_f32": (C.c_int, [_vp, _vp, _ll, _f, _vp, _vp]),
    "sei_gemm_bf16_atb": (C.c_int, [_vp, _vp, _vp, _ll, _i, _i, _ll, _ll, _vp]),
    "sei_gemm_bf16_tn_gelu_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _ll, _i, _i, _ll, _ll, _ll, _ll, _vp]),
    "sei_gemm_bf16_atb_accumulate": (C.c_int, [_vp, _vp, _vp, _ll, _i, _i, _ll, _ll, _vp]),
    "sei_gemm_bf16_tn": (C.c_int, [_vp, _vp, _vp, _vp, _ll, _i, _i, _ll, _ll, _ll, _i, _i, _vp]),
    "sei_gemm_bf16_tn_gelu_dual": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _ll, _ll, _ll, _vp]),
    "sei_gemm_bf16_tn_mul": (C.c_int, [_vp, _vp, _vp, _vp, _ll, _i, _i, _ll, _ll, _ll, _ll, _vp]),
    "sei_gemm_bf16_tn_rowscaled_bias": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i, _ll, _i, _i, _ll, _ll, _ll, _vp]),
    "sei_gemm_bf16_nn": (C.c_int, [_vp, _vp, _vp, _vp, _ll, _i, _i, _ll, _ll, _ll, _ll, _vp]),
    "sei_gemm_bf16_tn_residual": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _f, _ll, _i, _i, _ll, _ll, _ll, _ll, _vp]),
    "sei_ln_cl_forward_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _f, _vp]),
    "sei_ln_cl_backward_workspace_bytes": (C.c_longlong, [_i]),
    "sei_ln_cl_backward_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _vp]),
    "sei_colsum_bf16": (C.c_int, [_vp, _vp, _vp, _ll, _i, _vp]),
    "sei_conv3x3_igemm_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "sei_conv3x3_small_workspace_bytes": (C.c_longlong, [_i, _i]),
    "sei_conv3x3_small_forward_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "sei_conv3x3_small_backward_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "sei_dwconv7_workspace_bytes": (C.c_longlong, [_i]),
    "sei_dwconv7_cl_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "sei_dwconv7_cl_residual_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _f, _vp, _i, _i, _i, _i, _vp]),
    "sei_dwconv7_wgrad_cl_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "sei_gelu_bf16": (C.c_int, [_vp, _vp, _vp, _ll, _vp]),
    "sei_gelu_bwd_colsum_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _vp]),
    "sei_adam_step_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _vp]),
    "sei_transpose_bf16": (C.c_int, [_vp, _vp, _i, _i, _vp]),
    "sei_bias_pattern_add_bf16": (C.c_int, [_vp, _vp, _vp, _ll, _i, _i, _vp]),
    "sei_bias_pattern_grad_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _ll, _i, _i, _vp]),
    "sei_ln_small_workspace_bytes": (C.c_longlong, [_i]),
    "sei_ln_small_forward_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _f, _vp]),
    "sei_ln_small_backward_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _vp]),
    "sei_resize_bicubic_f32": (C.c_int, [_vp, _vp, _ll, _i, _i, _i, _i, _f, _f, _i, _vp]),
    "sei_resize_bicubic_backward_f32": (C.c_int, [_vp, _vp, _ll, _i, _i, _i, _i, _f, _f, _i, _vp]),
    "sei_rotate_nearest_f32": (C.c_int, [_vp, _vp, _ll, _i, _i, _vp, _vp]),
    "sei_rotate_nearest_backward_f32": (C.c_int, [_vp, _vp, _ll, _i, _i, _vp, _vp]),
    "sei_bgemm_tile_rows": (C.c_int, [_i, _i]),
    "sei_bgemm_bf16": (C.c_int, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _ll, _i, _ll, _ll, _i, _ll, _ll, _ll, _ll, _i, _ll, _ll,
                                 _vp]),
}


class SeiError(RuntimeError):
    pass


def load():
    """Load the CUDA library; raises if it has not been built (there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SeiError(
                f"{LIB_PATH} not found: the sm_100a CUDA library has not been built "
                "(run `python __graft_entry__.py build`). There is no CPU/PyTorch fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if lib.sei_abi_version() != 1:
            raise SeiError("libsei_b200.so ABI version mismatch; rebuild")
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load().sei_last_error().decode(errors="replace")
        raise SeiError(f"libsei_b200 call failed (code {rc}): {msg}")


def last_kernel():
    return load().sei_last_kernel().decode()


def launch_count():
    return int(load().sei_launch_count())
