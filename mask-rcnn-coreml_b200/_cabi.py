"""ctypes binding of libmaskrcnn_cuda.so (include/maskrcnn_cuda.h).

This is the only place the product touches native code.  There is no CPU
fallback: if the shared library is missing, or there is no sm_100 device,
loading / context creation raises.  Nothing here imports oracle/.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MRCNN_LIB_PATH: another build of the same library (A/B measurements of kernel changes on one box)
LIB_PATH = os.environ.get("MRCNN_LIB_PATH") or os.path.join(_HERE, "libmaskrcnn_cuda.so")

OK, EINVAL, ECUDA, ENCCL, EIO, ESTATE = 0, -1, -2, -3, -4, -5


class MaskRCNNError(RuntimeError):
    """Raised for every non-zero status (the Swift shim turns these into `throws`)."""

    def __init__(self, status, message):
        super().__init__(f"[mrcnn status {status}] {message}")
        self.status = status


class mrcnn_config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("device", C.c_int32),
        ("image_h", C.c_int32), ("image_w", C.c_int32),
        ("architecture", C.c_int32), ("num_classes", C.c_int32),
        ("bbox_std", C.c_float * 4),
        ("pre_nms_max_proposals", C.c_int32), ("max_proposals", C.c_int32),
        ("proposal_nms_iou", C.c_float),
        ("pool_size_classifier", C.c_int32), ("pool_size_mask", C.c_int32),
        ("fpn_selection_factor", C.c_float),
        ("max_detections", C.c_int32), ("detection_min_score", C.c_float),
        ("detection_nms_iou", C.c_float),
        ("mean_rgb", C.c_float * 3),
        ("max_batch", C.c_int32), ("precise_masks", C.c_int32),
        ("anchors_path", C.c_char_p), ("main_model_path", C.c_char_p),
        ("classifier_model_path", C.c_char_p), ("mask_model_path", C.c_char_p),
    ]


_vp, _i, _i64 = C.c_void_p, C.c_int, C.c_int64

# name -> (restype, argtypes); every symbol declared in include/maskrcnn_cuda.h
SIGNATURES = {
    "mrcnn_config_default": (None, [C.POINTER(mrcnn_config)]),
    "mrcnn_version": (C.c_char_p, []),
    "mrcnn_create": (_i, [C.POINTER(mrcnn_config), C.POINTER(_vp)]),
    "mrcnn_destroy": (None, [_vp]),
    "mrcnn_last_error": (C.c_char_p, [_vp]),
    "mrcnn_set_stream": (_i, [_vp, _vp]),
    "mrcnn_synchronize": (_i, [_vp]),
    "mrcnn_set_anchors": (_i, [_vp, _vp, _i64]),
    "mrcnn_set_weights": (_i, [_vp, _i, _vp, C.c_size_t]),
    "mrcnn_num_anchors": (_i64, [_vp]),
    "mrcnn_anchor_count": (_i64, [_i, _i]),
    "mrcnn_generate_anchors": (_i, [_i, _i, _vp, _i64]),
    "mrcnn_proposal_output_shape": (_i, [_vp, C.POINTER(_i64)]),
    "mrcnn_proposal_eval": (_i, [_vp, _i, _i64, _vp, _vp, _vp, _vp, _vp]),
    "mrcnn_pyramid_roialign_output_shape": (_i, [_vp, _i64, _i64, _i, C.POINTER(_i64)]),
    "mrcnn_pyramid_roialign_eval": (_i, [_vp, _i, _vp, _i, _i64, C.POINTER(_vp), C.POINTER(C.c_int32), _i64, _i, _vp, _vp]),
    "mrcnn_classifier_eval": (_i, [_vp, _i, _i64, _vp, _vp]),
    "mrcnn_classifier_select": (_i, [_vp, _i, _i64, _vp, _vp, _vp]),
    "mrcnn_detection_output_shape": (_i, [_vp, C.POINTER(_i64)]),
    "mrcnn_detection_eval": (_i, [_vp, _i, _i64, _vp, _vp, _vp, _vp, _vp]),
    "mrcnn_mask_eval": (_i, [_vp, _i, _i64, _vp, _vp, _vp]),
    "mrcnn_predict": (_i, [_vp, _i, _vp, _vp, _vp]),
    "mrcnn_detections_decode": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mrcnn_letterbox_eval": (_i, [_vp, _vp, _i, _i, _vp]),
    "mrcnn_letterbox_geometry": (_i, [_i, _i, _i, _i, C.POINTER(C.c_double)]),
    "mrcnn_unletterbox_boxes": (_i, [_i, _i, _i, _i, _vp, _i64, _i, _vp]),
    "mrcnn_nccl_unique_id": (_i, [_vp]),
    "mrcnn_comm_init": (_i, [_vp, _vp, _i, _i]),
    "mrcnn_predict_allgather": (_i, [_vp, _i, _vp, _vp, _vp]),
    "mrcnn_predict_submit": (_i, [_vp, _i, _vp, _vp, _vp, _i]),
    "mrcnn_predict_wait": (_i, [_vp]),
    "mrcnn_predict_in_flight": (_i, [_vp]),
    "mrcnn_last_stage_times": (_i, [_vp, _i, C.POINTER(C.c_char_p), C.POINTER(C.c_float)]),
    "mrcnn_launch_count": (_i64, [_vp]),
    "mrcnn_profile_enable": (_i, [_vp, _i]),
    "mrcnn_profile_read": (_i, [_vp, _i, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(_i64), C.POINTER(C.c_double)]),
    "mrcnn_roialign_nhwc_f16": (_i, [_vp, _i, _vp, _i, _i64, C.POINTER(_vp), C.POINTER(C.c_int32), _i64, _i, _vp, _vp]),
    "mrcnn_conv2d_nhwc_f16": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "mrcnn_debug_conv_trace": (_i, [_vp]),
    "mrcnn_debug_fused_expand_reduce": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp]),
    "mrcnn_backbone_eval": (_i, [_vp, _i, _vp, C.POINTER(_vp), _vp, _vp]),
}

_lib = None


def lib():
    """Loads libmaskrcnn_cuda.so (built in-tree by __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MaskRCNNError(ESTATE, f"{LIB_PATH} is missing: build it with "
                            "`make -C mask-rcnn-coreml_b200/csrc` (there is no CPU fallback)")
    l = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(l, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = l
    return l


_NP_NAMES = {"uint8": "uint8", "float32": "float32", "float16": "float16", "int32": "int32", "float64": "float64"}


def ptr(x, dtype=None, count=None):
    """Raw address of a numpy array (host) or a torch tensor (host or cuda); None -> NULL.
    dtype (a numpy dtype name, e.g. "float32") and count (minimum number of elements) are checked when given: the C ABI
    takes plain pointers, so a buffer of the wrong element type or size would be reinterpreted or overrun silently."""
    if x is None:
        return None
    if hasattr(x, "data_ptr"):            # torch.Tensor
        if not x.is_contiguous():
            raise MaskRCNNError(EINVAL, "tensor must be contiguous")
        if dtype is not None and str(x.dtype).replace("torch.", "") != _NP_NAMES[dtype]:
            raise MaskRCNNError(EINVAL, f"tensor must be {dtype}, got {x.dtype}")
        if count is not None and x.numel() < count:
            raise MaskRCNNError(EINVAL, f"tensor has {x.numel()} elements, {count} needed")
        if x.is_cuda:
            # The context's stream is non-blocking (include/maskrcnn_cuda.h, mrcnn_set_stream): whatever PyTorch has queued on
            # its current stream for this device (the kernels that produce this tensor, or still read it) must be done
            # before the library touches the buffer.  A host-side wait on an idle stream costs microseconds.
            import torch
            torch.cuda.current_stream(x.device).synchronize()
        return C.c_void_p(x.data_ptr())
    if hasattr(x, "ctypes"):              # numpy.ndarray
        if not x.flags["C_CONTIGUOUS"]:
            raise MaskRCNNError(EINVAL, "array must be C-contiguous")
        if dtype is not None and x.dtype.name != dtype:
            raise MaskRCNNError(EINVAL, f"array must be {dtype}, got {x.dtype}")
        if count is not None and x.size < count:
            raise MaskRCNNError(EINVAL, f"array has {x.size} elements, {count} needed")
        return C.c_void_p(x.ctypes.data)
    if isinstance(x, int):
        return C.c_void_p(x)
    raise MaskRCNNError(EINVAL, f"unsupported buffer type {type(x)}")


def check(ctx_handle, status):
    if status != OK:
        msg = lib().mrcnn_last_error(ctx_handle)
        raise MaskRCNNError(status, msg.decode() if msg else "unknown error")
