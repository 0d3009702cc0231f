"""Thin Python wrappers over single-operator entry points of libdbb200.so (tests and module drop-ins)."""
import torch

from . import _lib


def to_nhwc_bf16(x):
    """(N,C,H,W) float32 CUDA -> (N,H,W,C) bf16 CUDA through dbb_nchw_f32_to_nhwc_bf16."""
    _lib.require_cuda(x)
    x = x.contiguous().float()
    n, c, h, w = x.shape
    y = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().dbb_nchw_f32_to_nhwc_bf16(x.data_ptr(), y.data_ptr(), n, c, h, w, _lib.stream_ptr()), "nchw->nhwc")
    return y


def to_nchw_f32(x):
    """(N,H,W,C) bf16 CUDA -> (N,C,H,W) float32."""
    _lib.require_cuda(x)
    x = x.contiguous()
    n, h, w, c = x.shape
    y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().dbb_nhwc_bf16_to_nchw_f32(x.data_ptr(), y.data_ptr(), n, c, h, w, _lib.stream_ptr()), "nhwc->nchw")
    return y


def conv2d_raw(kind, x_nhwc, weight, bias, n, h, w, cin, cout, ks, stride, pad, out_shape):
    """kind 0 Conv2d fprop, 1 Conv2d dgrad, 2 ConvT(k2,s2) fprop, 3 ConvT dgrad. (h, w) = forward-input extent."""
    L = _lib.lib()
    _lib.require_cuda(x_nhwc, weight)
    y = torch.empty(out_shape, dtype=torch.bfloat16, device=x_nhwc.device)
    wsb = L.dbb_conv2d_workspace(kind, n, h, w, cin, cout, ks, stride, pad)
    ws = torch.empty(wsb, dtype=torch.uint8, device=x_nhwc.device)
    weight = weight.contiguous().float()
    bptr = bias.contiguous().float().data_ptr() if bias is not None else None
    with torch.cuda.device(x_nhwc.device):
        _lib.check(L.dbb_conv2d(kind, x_nhwc.data_ptr(), weight.data_ptr(), bptr, y.data_ptr(), n, h, w, cin, cout, ks,
                                stride, pad, ws.data_ptr(), wsb, _lib.stream_ptr()), "dbb_conv2d")
    return y


def conv2d_wgrad_raw(kind, x_nhwc, dy_nhwc, n, h, w, cin, cout, ks, stride, pad):
    L = _lib.lib()
    _lib.require_cuda(x_nhwc, dy_nhwc)
    shape = (cout, cin, ks, ks) if kind == 0 else (cin, cout, 2, 2)
    dw = torch.empty(shape, dtype=torch.float32, device=x_nhwc.device)
    with torch.cuda.device(x_nhwc.device):
        _lib.check(L.dbb_conv2d_wgrad(kind, x_nhwc.data_ptr(), dy_nhwc.data_ptr(), dw.data_ptr(), n, h, w, cin, cout, ks,
                                      stride, pad, None, 0, _lib.stream_ptr()), "dbb_conv2d_wgrad")
    return dw
