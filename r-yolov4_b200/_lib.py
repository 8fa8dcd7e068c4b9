"""ctypes binding of libryolo_b200.so (the C ABI declared in include/ryolo_b200.h).

There is no CPU fallback: if the shared library is missing or a kernel reports an error the call
raises.  torch is used only as the owner of device memory and streams.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libryolo_b200.so")

_vp, _i64, _i32, _f32, _sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/ryolo_b200.h
SIGNATURES = {
    "ryolo_abi_version": (_i32, []),
    "ryolo_last_error": (ctypes.c_char_p, []),
    "ryolo_set_error": (None, [ctypes.c_char_p]),
    "ryolo_check_device": (_i32, [_i32]),
    "ryolo_encode_labels": (_i32, [_vp, ctypes.c_longlong, _i32, _vp, _vp]),
    "ryolo_tune": (_i32, [ctypes.c_char_p, _i32]),
    "ryolo_knob": (_i32, [_i32]),
    "ryolo_pairwise_iou_rotated_workspace": (_sz, [_i64, _i64]),
    "ryolo_pairwise_iou_rotated": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp, _sz, _vp]),
    "ryolo_eval_match_workspace": (_sz, [_i64, _i32]),
    "ryolo_eval_match": (_i32, [_vp, _vp, _i64, _i32, _vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _sz, _vp]),
    "ryolo_nms_rotated_workspace": (_sz, [_i64]),
    "ryolo_nms_rotated": (_i32, [_vp, _vp, _i64, _f32, _vp, _vp, _vp, _sz, _vp]),
    "ryolo_post_process_workspace": (_sz, [_i64, _i64, _i32, _i32]),
    "ryolo_post_process": (_i32, [_vp, _i64, _i64, _i32, _f32, _f32, _i32, _i32, _f32, _i32, _vp, _vp, _vp, _vp,
                                  _sz, _vp]),
    "ryolo_decode_csl": (_i32, [_vp, _i64, _i32, _i32, _f32, ctypes.POINTER(_f32), _vp, _i64, _i64, _vp]),
    "ryolo_decode_kfiou": (_i32, [_vp, _i64, _i32, _i32, _i32, _f32, _vp, _vp, _i64, _i64, _vp]),
    "ryolo_kfloss": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "ryolo_pos_record_bytes": (_sz, []),
    "ryolo_loss_workspace": (_sz, [_i64, _i32, ctypes.POINTER(ctypes.c_int32), _i64]),
    "ryolo_build_targets": (_i32, [_vp, _i64, _i32, _i32, _vp, _i32, _i64, ctypes.POINTER(ctypes.c_int32), _vp, _vp,
                                   _vp, _sz, _vp]),
    "ryolo_loss": (_i32, [_i32, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i64, _i32, _i32,
                          ctypes.POINTER(ctypes.c_int32), _vp, _i64, _i32, _vp, ctypes.POINTER(_f32), _vp, _vp, _sz,
                          _vp]),
}

_f32p, _dbl, _ll = ctypes.POINTER(_f32), ctypes.c_double, ctypes.c_longlong


class BnFuse(ctypes.Structure):
    """struct ryolo_bn_fuse (include/ryolo_b200.h)."""
    _fields_ = [("partial", _vp), ("sum", _vp), ("sumsq", _vp), ("counter", _vp), ("gamma", _vp), ("beta", _vp), ("running_mean", _vp),
                ("running_var", _vp), ("num_batches", _vp), ("eps", _f32), ("momentum", _f32), ("scale", _vp),
                ("shift", _vp), ("save_mean", _vp), ("save_invstd", _vp)]


class ConvDesc(ctypes.Structure):
    """struct ryolo_conv_desc (include/ryolo_b200.h)."""
    _fields_ = [("x", _vp), ("N", _i32), ("H", _i32), ("W", _i32), ("Cin", _i32), ("x_cpitch", _ll), ("w", _vp),
                ("Cout", _i32), ("ksize", _i32), ("stride", _i32), ("out", _vp), ("out_mode", _i32),
                ("out_cpitch", _ll), ("scale", _vp), ("shift", _vp), ("act", _i32), ("residual", _vp),
                ("res_cpitch", _ll), ("head_na", _i32), ("head_ch", _i32), ("bn", ctypes.POINTER(BnFuse))]


SIGNATURES.update({
    "ryolo_conv2d_forward": (_i32, [ctypes.POINTER(ConvDesc), _vp]),
    "ryolo_conv2d_reference": (_i32, [ctypes.POINTER(ConvDesc), _vp]),
    "ryolo_bn_stats": (_i32, [_vp, _ll, _ll, _i32, _vp, _vp, _vp]),
    "ryolo_bn_finalize": (_i32, [_vp, _vp, _dbl, _i32, _vp, _vp, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ryolo_scale_shift_act": (_i32, [_vp, _ll, _vp, _vp, _vp, _ll, _vp, _vp, _i32, _vp, _ll, _vp, _ll, _ll, _i32, _vp]),
    "ryolo_scale_shift_act_bn": (_i32, [_vp, _ll, ctypes.POINTER(BnFuse), ctypes.c_double, _i32, _vp, _ll, _vp, _ll, _ll,
                                        _i32, _vp]),
    "ryolo_maxpool": (_i32, [_vp, _ll, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _ll, _vp]),
    "ryolo_resize_copy": (_i32, [_vp, _ll, _i32, _i32, _i32, _i32, _i32, _vp, _ll, _vp]),
    "ryolo_stem_im2col": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "ryolo_pack_weights": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "ryolo_pack_weights_multi": (_i32, [_vp, _i32, _ll, _vp]),
    "ryolo_conv2d_dgrad": (_i32, [_vp, _ll, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _ll, _i32, _vp]),
    "ryolo_conv2d_wgrad": (_i32, [_vp, _ll, _i32, _i32, _i32, _i32, _vp, _ll, _i32, _i32, _i32, _i32, _vp, _vp]),
    "ryolo_unpack_wgrad_multi": (_i32, [_vp, _i32, _ll, _vp]),
    "ryolo_bn_act_bwd": (_i32, [_vp, _ll, _vp, _ll, _vp, _vp, _vp, _vp, _i32, _ll, _i32, _vp, _vp, _ll, _vp, _vp, _vp, _ll, _vp]),
    "ryolo_act_bwd2": (_i32, [_vp, _ll, _vp, _ll, _vp, _vp, _vp, _ll, _vp, _vp, _i32, _vp, _ll, _ll, _i32, _vp]),
    "ryolo_add_into": (_i32, [_vp, _ll, _vp, _ll, _ll, _i32, _i32, _vp]),
    "ryolo_maxpool_bwd": (_i32, [_vp, _ll, _vp, _ll, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _ll, _i32, _vp,
                                 _vp]),
    "ryolo_spp_bwd": (_i32, [_vp, _ll, _vp, _vp, _vp, _ll, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _ll, _i32, _vp]),
    "ryolo_upsample2x_bwd": (_i32, [_vp, _ll, _i32, _i32, _i32, _i32, _vp, _ll, _i32, _vp]),
    "ryolo_head_grad_pack": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ryolo_sgd_step": (_i32, [_vp, _vp, _vp, _ll, _f32, _f32, _f32, _i32, _i32, _vp]),
    "ryolo_sgd_step_lrdev": (_i32, [_vp, _vp, _vp, _ll, _vp, _f32, _f32, _i32, _vp]),
    "ryolo_adam_step": (_i32, [_vp, _vp, _vp, _vp, _ll, _f32, _f32, _f32, _f32, _f32, _i32, _vp]),
})

_lib = None


class RyoloError(RuntimeError):
    pass


def lib():
    """Load (once) and return the CDLL.  Raises if the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RyoloError(
                f"{SO_PATH} not found: build it with `python r-yolov4_b200/build.py` "
                "(there is no CPU / PyTorch fallback for the hot path)")
        h = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name, None)
            if fn is None:
                continue
            fn.restype, fn.argtypes = res, args
        _lib = h
    return _lib


LAUNCHES = [0]          # kernels launched through the library (bench.py reports it as gpu_launches)


def count(n):
    LAUNCHES[0] += n


def check(rc):
    if rc != 0:
        raise RyoloError(f"libryolo_b200 error {rc}: {lib().ryolo_last_error().decode()}")


def tune(**kw):
    """Set process-wide tuning knobs of the library (ryolo_tune), e.g. tune(halo=1, wg_split=0)."""
    for k, v in kw.items():
        check(lib().ryolo_tune(k.encode(), int(v)))


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, what):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RyoloError(f"{what} must be a CUDA tensor: the ryolo_b200 hot path has no CPU fallback")


_ws = {}


def workspace(nbytes, device, tag="default"):
    """Grow-only scratch buffer per (device, tag)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), tag)
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf
