"""argtypes/restype declarations for the C ABI beyond m3d_conv2d_nhwc."""
import ctypes as C

vp, i, f, sz, lg, db = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_long, C.c_double

_SIGS = {
    "m3d_dcn_v2_forward": [vp, vp, vp, vp, vp, vp] + [i] * 15 + [vp, sz, vp],
    "m3d_dcn_v2_backward": [vp] * 10 + [i] * 15 + [vp, sz, vp],
    "m3d_nms": [vp, vp, vp, i, i, f, i],
    "m3d_nms_batched": [vp, i, vp, i, i, f, vp, sz, vp, vp, vp],
    "m3d_decode_topk": [vp, vp, vp, vp, vp, vp, vp, i, i, i, i, f, f, i, vp, vp, vp, vp, sz, vp],
    "m3d_decode_topk_heads": [vp, vp, vp, i, vp, vp, vp, vp, i, i, i, i, f, f, i, vp, vp, vp, vp, sz, vp],
    "m3d_gather_kept": [vp, i, i, i, vp, vp, i, vp, vp],
    "m3d_stem_conv7x7": [vp, vp, vp, vp, i, i, i, i, i, f, vp],
    "m3d_conv2d_wgrad": [vp, i, i, vp, i, i, vp] + [i] * 12 + [vp, sz, vp],
    "m3d_channel_sum": [vp, lg, i, i, i, vp, vp, sz, vp],
    "m3d_preprocess_u8": [vp, vp, i, i, i, vp, vp, i, vp],
    "m3d_preprocess_u8_pad": [vp, vp, vp, vp, vp, i, i, i, vp, vp, i, vp],
    "m3d_stem_conv7x7_s2d": [vp, vp, vp, vp, i, i, i, f, vp],
    "m3d_maxpool2x2_nhwc": [vp, vp, i, i, i, i, i, i, i, vp],
    "m3d_upsample_add_nhwc": [vp, vp, vp, vp, i, i, i, i, i, i, i, i, i, vp],
    "m3d_upsample_backward": [vp, vp, vp, vp, vp, i, i, i, i, i, vp, sz, vp],
    "m3d_cls_softmax": [vp, i, i, i, i, i, i, vp, vp, vp, vp, vp, vp, vp],
    "m3d_cls_softmax_shape_om": [vp, i, i, i, i, i, i, vp, vp, vp, vp, vp, vp, vp, i, f, f, vp, vp],
    "m3d_center_align_om2": [vp, vp, vp, i, vp, vp, vp, vp, i, f, f, vp, vp, i, lg, vp],
    "m3d_shape_align_om": [vp, vp, vp, i, f, f, vp, lg, vp],
    "m3d_center_align_om": [vp, vp, vp, i, i, i, vp, i, f, f, f, f, f, f, vp, i, lg, vp],
    "m3d_set_sm_limit": [i],
    "m3d_set_pdl": [i],
    "m3d_compute_targets": [vp, vp, vp, vp, i, vp, vp, i, vp, i, i, i, i, f, db, db, db, db, db, vp, vp, vp, vp, vp, vp, vp, vp,
                            vp, vp, sz, vp],
    "m3d_head_mlp": [vp, i, i, i, i, i, i, vp, vp, vp, vp, vp, vp, i, i, i, vp, i, i, f, vp],
    "m3d_refine_3d": [vp, vp, i, i, i, vp, vp, f, i, db, db, vp, vp, vp],
    "m3d_flatten_heads": [vp, i, i, i, i, i, vp, vp, vp, vp],
    "m3d_anab_pool": [vp, i, i, i, i, i, i, i, vp, vp, sz, vp, vp, vp],
    "m3d_anab_attention": [vp, i, vp, vp, vp, i, i, vp, vp, f, vp, i, i, i, i, i, i, vp, sz, vp],
    "m3d_nchw_to_nhwc": [vp, i, vp, i, i, i, i, i, i, i, vp],
    "m3d_nhwc_to_nchw": [vp, i, vp, i, i, i, i, i, i, i, vp],
}
_SIZE_FNS = {
    "m3d_conv_desc_size": [],
    "m3d_dcn_v2_forward_workspace": [i] * 12,
    "m3d_dcn_v2_backward_workspace": [i] * 11,
    "m3d_nms_workspace_bytes": [i, i],
    "m3d_anab_pool_workspace": [i, i, i, vp, i, i],
    "m3d_anab_attention_workspace": [i, i],
    "m3d_decode_topk_workspace": [i],
    "m3d_compute_targets_workspace": [i, i],
    "m3d_conv2d_wgrad_workspace": [i] * 7,
    "m3d_channel_sum_workspace": [i],
    "m3d_upsample_backward_workspace": [i, i],
}


def declare(L):
    for name, args in _SIGS.items():
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = args
    for name, args in _SIZE_FNS.items():
        fn = getattr(L, name)
        fn.restype = C.c_size_t
        fn.argtypes = args


def exported_names():
    return ["m3d_last_error", "m3d_last_kernel", "m3d_version", "m3d_conv2d_nhwc"] + list(_SIGS) + list(_SIZE_FNS)
