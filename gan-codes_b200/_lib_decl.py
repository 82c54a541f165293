"""argtypes/restype declarations for the entry points beyond the two GEMM engines (kept separate so
``_lib.py`` stays a readable mirror of the first half of include/dfgan_b200.h)."""
from __future__ import annotations

import ctypes as C

P, I, L, F = C.c_void_p, C.c_int, C.c_int64, C.c_float


class LinearItem(C.Structure):
    _fields_ = [("inp", P), ("in_stride", L), ("w", P), ("bias", P), ("out", P), ("out_stride", L),
                ("mask", P), ("k", C.c_int32), ("n_out", C.c_int32), ("o0", C.c_int32), ("relu", C.c_int32)]


class AugItem(C.Structure):
    """struct dfgan_aug_item of include/dfgan_b200.h."""
    _fields_ = [("src", P), ("pitch", L)] + [(n, C.c_int32) for n in (
        "src_w", "src_h", "x1", "y1", "cw", "ch", "rw", "rh", "top", "left", "flip", "reserved")]


def declare(l: C.CDLL) -> None:
    l.dfgan_pack_entry_size.argtypes = []
    l.dfgan_pack_entry_size.restype = C.c_int
    l.dfgan_pack_entry_bs_offset.argtypes = []
    l.dfgan_pack_entry_bs_offset.restype = C.c_int
    l.dfgan_unpack_entry_size.argtypes = []
    l.dfgan_unpack_entry_size.restype = C.c_int
    l.dfgan_augment_workspace_bytes.argtypes = [I, I]
    l.dfgan_augment_workspace_bytes.restype = C.c_size_t
    l.dfgan_text_encoder_workspace_bytes.argtypes = [I, I]
    l.dfgan_text_encoder_workspace_bytes.restype = C.c_size_t
    l.dfgan_peer_flag_bytes.argtypes = []
    l.dfgan_peer_flag_bytes.restype = C.c_size_t
    l.dfgan_peer_handle_bytes.argtypes = []
    l.dfgan_peer_handle_bytes.restype = C.c_int
    for name, argtypes in _TABLE.items():
        fn = getattr(l, name)   # AttributeError = header and library disagree: fail loudly
        fn.argtypes = argtypes
        fn.restype = C.c_int


_TABLE: dict = {
    "dfgan_peer_alloc": [C.c_size_t, P],
    "dfgan_peer_free": [P],
    "dfgan_peer_export": [P, P],
    "dfgan_peer_import": [P, P],
    "dfgan_peer_close": [P],
    "dfgan_peer_allreduce_f32": [P, P, I, I, L, I, I, I, I, P],
    "dfgan_peer_allreduce_mc_f32": [P, P, I, I, L, I, I, I, I, P],
    "dfgan_peer_reduce_scatter_f32": [P, P, I, I, L, I, I, I, I, P],
    "dfgan_peer_allgather_f32": [P, P, I, I, L, I, I, I, P],
    "dfgan_peer_sum_scalar_f32": [P, P, P, P, I, I, I, I, P],
    "dfgan_affine2_fwd": [P, P, P, P, P, P, L, I, I, I, P],
    "dfgan_affine2_bwd": [P, P, P, P, P, P, P, P, P, P, P, P, L, P, P, P, I, I, I, P],
    "dfgan_pool": [P, P, P, P, F, I, I, I, I, I, F, P],
    "dfgan_dres_bwd": [P, P, P, P, P, P, I, I, I, P],
    "dfgan_dot": [P, P, L, P, F, P],
    "dfgan_im2col3": [P, P, P, I, I, I, P],
    "dfgan_stem_fwd": [P, P, P, P, P, P, I, I, I, P],
    "dfgan_tanh_bwd_pack": [P, P, P, P, I, I, P],
    "dfgan_nchw_to_nhwc": [P, P, I, I, I, P],
    "dfgan_nhwc_to_nchw": [P, P, I, I, I, I, P],
    "dfgan_rowsq": [P, I, L, P, P],
    "dfgan_head_cat": [P, P, P, P, I, I, I, P],
    "dfgan_head_cat_bwd": [P, P, P, I, I, I, P],
    "dfgan_head_out": [P, P, P, I, I, P],
    "dfgan_head_out_bwd": [P, P, P, P, P, P, I, I, P],
    "dfgan_linear_batched": [P, I, I, I, I, I, P],
    "dfgan_pack_weights": [P, P, I, I, I, I, I, I, I, P, F, I, P],
    "dfgan_pack_entry_fill": [P, P, P, I, I, I, I, I, I, I, P, F, I, I],
    "dfgan_pack_batched": [P, I, I, P],
    "dfgan_unfold_subpixel_wgrad": [P, P, I, I, P],
    "dfgan_unpack_entry_fill": [P, P, P, I, I, I, I],
    "dfgan_wgrad_unpack": [P, I, I, P],
    "dfgan_sqnorm": [P, L, P, P],
    "dfgan_sqnorm_det": [P, L, P, P, I, P],
    "dfgan_adam_step": [P, P, P, P, L, P, P, F, F, F, F, F, F, P],
    "dfgan_gblock_small_grads": [P, P, P, P, P, P, I, P],
    "dfgan_d_hinge": [P, P, P, I, P],
    "dfgan_gp_coef": [P, P, P, I, P],
    "dfgan_txtimg_loss": [P, I, I, L, P, I, P, I, P, P],
    "dfgan_conv2d_f32_fprop": [P, P, P, P, I, I, I, I, I, I, I, I, I, I, I, P],
    "dfgan_conv2d_f32_dgrad": [P, P, P, I, I, I, I, I, I, I, I, I, I, I, P],
    "dfgan_conv2d_f32_wgrad": [P, P, P, I, I, I, I, I, I, I, I, I, I, I, P],
    "dfgan_augment_check": [P, I],
    "dfgan_augment": [P, P, P, P, P, C.c_size_t, I, I, P],
    "dfgan_text_encoder_fwd": [P, P, P, L, P, P, P, P, P, P, C.c_size_t, I, I, I, I, P],
}

EXPORTS = ["dfgan_conv_run", "dfgan_wgrad_run", "dfgan_version", "dfgan_arch", "dfgan_launch_count",
           "dfgan_profile_enable", "dfgan_profile_read", "dfgan_pack_entry_size", "dfgan_pack_entry_bs_offset", "dfgan_conv_uses_image",
           "dfgan_unpack_entry_size", "dfgan_conv_splitk_bytes", "dfgan_peer_flag_bytes", "dfgan_peer_handle_bytes", "dfgan_text_encoder_workspace_bytes", "dfgan_augment_workspace_bytes",
           *_TABLE.keys()]
