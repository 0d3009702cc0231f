"""ctypes binding of libdbb200.so (include/dbb200.h).  Fails loudly: there is no CPU fallback."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdbb200.so")


class DbbLossState(C.Structure):
    _fields_ = [("sums", C.c_double * 12), ("n_pos", C.c_longlong), ("n_neg", C.c_longlong),
                ("n_above", C.c_longlong), ("n_tie", C.c_longlong), ("tau", C.c_float),
                ("tau_bits", C.c_uint), ("tie_ticket", C.c_int), ("reduction", C.c_int), ("coef", C.c_float * 8)]


class DbbCandidate(C.Structure):
    _fields_ = [("kind", C.c_int), ("first_y", C.c_int), ("first_x", C.c_int),
                ("x0", C.c_int), ("y0", C.c_int), ("x1", C.c_int), ("y1", C.c_int), ("count", C.c_int),
                ("sum", C.c_double), ("keep", C.c_int), ("pad_", C.c_int)]


_lib = None


class DbbError(RuntimeError):
    pass


def lib():
    """Loads the CUDA library; raises if it has not been built (python -m db_text_minimal_b200._build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DbbError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                       "(nvcc, sm_100a). db_text_minimal_b200 has no CPU or PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, f32, f64, sz = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_double, C.c_size_t
    L.dbb_version.restype = i32
    L.dbb_strerror.restype = C.c_char_p
    L.dbb_strerror.argtypes = [i32]
    L.dbb_last_cuda_error.restype = C.c_char_p
    L.dbb_launch_count.restype = C.c_uint64

    def sig(name, restype, argtypes):
        fn = getattr(L, name, None)
        if fn is not None:
            fn.restype = restype
            fn.argtypes = argtypes

    sig("dbb_dbloss_workspace", sz, [i64, i32, i64, i64, i32])
    sig("dbb_dbloss_fwd", i32, [vp, vp, i64, i32, i64, i64, f32, f32, i32, f32, f32, vp, vp, vp, sz, vp])
    sig("dbb_dbloss_bwd", i32, [vp, vp, i64, i32, i64, i64, f32, f32, i32, f32, vp, vp, vp, vp])
    sig("dbb_conv1_workspace", sz, [i64, i64, i64, i32])
    sig("dbb_conv1_fwd", i32, [vp, vp, vp, i64, i64, i64, vp, sz, vp])
    sig("dbb_conv1_wgrad", i32, [vp, vp, vp, i64, i64, i64, vp, sz, vp])
    sig("dbb_thresh_map", i32, [vp, i64, i64, i64, vp, vp, vp, vp, vp, i32, i32, vp])
    sig("dbb_adam_step", i32, [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, f32, vp, vp])
    sig("dbb_text_score_hist", i32, [vp, i64, vp, vp, i64, i64, i64, f32, vp, vp])
    sig("dbb_step_fwd", i32, [vp, vp, vp, i64, f32, vp])
    sig("dbb_step_bwd", i32, [vp, vp, vp, vp, vp, i64, f32, vp])
    sig("dbb_postprocess_workspace", sz, [i64, i64, i64])
    sig("dbb_binarize", i32, [vp, i64, i32, i64, i64, f32, vp, vp])
    sig("dbb_binarize_ccl_score", i32, [vp, i64, i32, i64, i64, f32, f64, vp, vp, vp, vp, i32, vp, sz, vp])
    sig("dbb_ccl_border_points", i32, [vp, sz, i64, i64, i64, vp, vp, i32, vp])
    sig("dbb_boxes_from_border_points", i32, [vp, vp, vp, vp, i64, i32, i32, i64, i64, vp, f32, i32, vp, vp, vp, vp, i32])
    sig("dbb_mini_box", i32, [vp, i32, vp, vp])
    sig("dbb_polygons_from_bitmap", i32, [vp, vp, vp, i64, i32, i64, i64, vp, f32, i32, vp, vp, i32, vp, vp, i32])
    sig("dbb_trace_contour", i32, [vp, i64, i64, i32, i32, i32, vp, i32])
    sig("dbb_approx_poly_dp", i32, [vp, i32, f64, vp, i32, vp])
    sig("dbb_fill_polygons", i32, [vp, vp, vp, vp, i32, vp, i64, i64, vp])
    sig("dbb_unpack_batch", i32, [vp, f32, f32, f32, vp, vp, vp, vp, i64, i64, i64, vp, vp, vp])
    sig("dbb_clipper_offset", i32, [vp, i32, f64, f64, vp, i32, vp, i32])
    sig("dbb_clipper_offset_raw", i32, [vp, i32, f64, f64, vp, i32])
    sig("dbb_net_num_params", i32, [])
    sig("dbb_net_param_name", C.c_char_p, [i32])
    sig("dbb_net_param_numel", i32, [i32])
    sig("dbb_net_num_buffers", i32, [])
    sig("dbb_net_buffer_name", C.c_char_p, [i32])
    sig("dbb_net_buffer_numel", i32, [i32])
    sig("dbb_net_create", vp, [i64, i64, i64, i32])
    sig("dbb_net_create_ex", vp, [i64, i64, i64, i32, i32])
    sig("dbb_net_precision", i32, [vp])
    sig("dbb_net_backward_ex", i32, [vp, vp, vp, vp, vp, vp, vp, sz, i32, vp])
    sig("dbb_net_destroy", None, [vp])
    sig("dbb_net_workspace_bytes", sz, [vp])
    sig("dbb_net_out_channels", i64, [vp])
    sig("dbb_net_flops_fwd", C.c_uint64, [vp])
    sig("dbb_net_forward", i32, [vp, vp, vp, vp, vp, vp, sz, vp])
    sig("dbb_net_num_segments", i32, [])
    sig("dbb_net_backward", i32, [vp, vp, vp, vp, vp, vp, sz, i32, vp])
    sig("dbb_profile_enable", None, [i32])
    sig("dbb_profile_report", sz, [C.c_char_p, sz])
    sig("dbb_ops_workspace", sz, [])
    sig("dbb_bn_fwd", i32, [vp, i64, i32, vp, vp, vp, vp, i32, vp, i32, vp, vp, vp, sz, vp])
    sig("dbb_bn_bwd", i32, [vp, vp, vp, i64, i32, vp, vp, vp, vp, vp, vp, vp, sz, vp])
    sig("dbb_maxpool_fwd", i32, [vp, i64, i64, i64, i32, vp, vp, vp])
    sig("dbb_maxpool_bwd", i32, [vp, vp, i64, i64, i64, i32, vp, vp])
    sig("dbb_upsample_add_fwd", i32, [vp, i64, i64, vp, i64, i64, i64, i32, vp, vp])
    sig("dbb_upsample_into", i32, [vp, i64, i64, i64, i64, i64, i32, vp, i32, i32, vp])
    sig("dbb_upsample_bwd", i32, [vp, i32, i32, i64, i64, i64, i32, vp, i64, i64, i32, vp])
    sig("dbb_head_tail_fwd", i32, [vp, i64, i64, i64, vp, vp, vp, vp, i32, vp, vp, vp, vp, f32, i32, vp, vp, vp, sz, vp])
    sig("dbb_head_tail_bwd", i32, [vp, i64, i64, i64, vp, vp, vp, vp, vp, vp, f32, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp])
    sig("dbb_net_debug_shape", i32, [vp, C.c_char_p, vp])
    sig("dbb_net_debug_read", i32, [vp, C.c_char_p, vp, vp, vp])
    sig("dbb_conv2d", i32, [i32, vp, vp, vp, vp, i64, i64, i64, i32, i32, i32, i32, i32, vp, sz, vp])
    sig("dbb_conv2d_workspace", sz, [i32, i64, i64, i64, i32, i32, i32, i32, i32])
    sig("dbb_conv2d_wgrad_workspace", sz, [])
    sig("dbb_conv2d_wgrad", i32, [i32, vp, vp, vp, i64, i64, i64, i32, i32, i32, i32, i32, vp, sz, vp])
    sig("dbb_nchw_f32_to_nhwc_bf16", i32, [vp, vp, i64, i32, i64, i64, vp])
    sig("dbb_nhwc_bf16_to_nchw_f32", i32, [vp, vp, i64, i32, i64, i64, vp])
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        L = lib()
        raise DbbError(f"{what}: {L.dbb_strerror(rc).decode()} [{L.dbb_last_cuda_error().decode()}]")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise DbbError("db_text_minimal_b200 runs on CUDA (sm_100a) tensors only; got a %s tensor. "
                           "There is no CPU fallback." % t.device)


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def profile_enable(on=True):
    lib().dbb_profile_enable(1 if on else 0)


def profile_report():
    """[{name, launches, ms}] of the launches recorded since profile_enable(); synchronises the device."""
    import json
    buf = C.create_string_buffer(1 << 20)
    n = lib().dbb_profile_report(buf, len(buf))
    return json.loads(buf.value.decode())["kernels"] if n else []
