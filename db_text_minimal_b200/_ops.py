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
    wsb = L.dbb_conv2d_wgrad_workspace()
    ws = torch.empty(wsb, dtype=torch.uint8, device=x_nhwc.device)
    with torch.cuda.device(x_nhwc.device):
        _lib.check(L.dbb_conv2d_wgrad(kind, x_nhwc.data_ptr(), dy_nhwc.data_ptr(), dw.data_ptr(), n, h, w, cin, cout, ks,
                                      stride, pad, ws.data_ptr(), wsb, _lib.stream_ptr()), "dbb_conv2d_wgrad")
    return dw


def _ws(device):
    L = _lib.lib()
    n = L.dbb_ops_workspace()
    return torch.empty(n, dtype=torch.uint8, device=device), n


def bn_fwd(z_nhwc, gamma, beta, running_mean=None, running_var=None, training=True, residual=None, relu=True):
    """z (N,H,W,C) bf16 -> (out bf16, stats4 float32[4C])."""
    L = _lib.lib()
    _lib.require_cuda(z_nhwc, gamma, beta)
    c = z_nhwc.shape[-1]
    px = z_nhwc.numel() // c
    out = torch.empty_like(z_nhwc)
    stats = torch.empty(4 * c, dtype=torch.float32, device=z_nhwc.device)
    ws, n = _ws(z_nhwc.device)
    with torch.cuda.device(z_nhwc.device):
        _lib.check(L.dbb_bn_fwd(z_nhwc.data_ptr(), px, c, gamma.data_ptr(), beta.data_ptr(),
                                running_mean.data_ptr() if running_mean is not None else None,
                                running_var.data_ptr() if running_var is not None else None, int(training),
                                residual.data_ptr() if residual is not None else None, int(relu), out.data_ptr(),
                                stats.data_ptr(), ws.data_ptr(), n, _lib.stream_ptr()), "dbb_bn_fwd")
    return out, stats


def bn_bwd(dout, act, z, gamma, stats, want_dres=False):
    L = _lib.lib()
    c = z.shape[-1]
    px = z.numel() // c
    dz = torch.empty_like(z)
    dres = torch.empty_like(z) if want_dres else None
    dgamma = torch.empty(c, dtype=torch.float32, device=z.device)
    dbeta = torch.empty(c, dtype=torch.float32, device=z.device)
    ws, n = _ws(z.device)
    with torch.cuda.device(z.device):
        _lib.check(L.dbb_bn_bwd(dout.data_ptr(), act.data_ptr() if act is not None else None, z.data_ptr(), px, c,
                                gamma.data_ptr(), stats.data_ptr(), dz.data_ptr(), dres.data_ptr() if want_dres else None,
                                dgamma.data_ptr(), dbeta.data_ptr(), ws.data_ptr(), n, _lib.stream_ptr()), "dbb_bn_bwd")
    return dz, dres, dgamma, dbeta


def maxpool_fwd(x):
    L = _lib.lib()
    n, h, w, c = x.shape
    oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    y = torch.empty((n, oh, ow, c), dtype=torch.bfloat16, device=x.device)
    am = torch.empty((n, oh, ow, c), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(L.dbb_maxpool_fwd(x.data_ptr(), n, h, w, c, y.data_ptr(), am.data_ptr(), _lib.stream_ptr()), "maxpool_fwd")
    return y, am


def maxpool_bwd(dy, am, h, w):
    L = _lib.lib()
    n, oh, ow, c = dy.shape
    dx = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=dy.device)
    with torch.cuda.device(dy.device):
        _lib.check(L.dbb_maxpool_bwd(dy.data_ptr(), am.data_ptr(), n, h, w, c, dx.data_ptr(), _lib.stream_ptr()), "maxpool_bwd")
    return dx


def upsample_add(xs, y):
    L = _lib.lib()
    n, hs, ws_, c = xs.shape
    _, h, w, _ = y.shape
    out = torch.empty_like(y)
    with torch.cuda.device(y.device):
        _lib.check(L.dbb_upsample_add_fwd(xs.data_ptr(), hs, ws_, y.data_ptr(), n, h, w, c, out.data_ptr(), _lib.stream_ptr()), "upsample_add")
    return out


def upsample_into(xs, dst, coff):
    L = _lib.lib()
    n, hs, ws_, c = xs.shape
    _, h, w, ct = dst.shape
    with torch.cuda.device(dst.device):
        _lib.check(L.dbb_upsample_into(xs.data_ptr(), hs, ws_, n, h, w, c, dst.data_ptr(), ct, coff, _lib.stream_ptr()), "upsample_into")
    return dst


def upsample_bwd(d_big, coff, c, hs, ws_, d_xs=None):
    L = _lib.lib()
    n, h, w, ct = d_big.shape
    acc = d_xs is not None
    if d_xs is None:
        d_xs = torch.empty((n, hs, ws_, c), dtype=torch.bfloat16, device=d_big.device)
    with torch.cuda.device(d_big.device):
        _lib.check(L.dbb_upsample_bwd(d_big.data_ptr(), ct, coff, n, h, w, c, d_xs.data_ptr(), hs, ws_, int(acc), _lib.stream_ptr()), "upsample_bwd")
    return d_xs


def head_tail_fwd(zt, gamma, beta, rm, rv, training, w2b, w2t, b2b, b2t, k=50.0):
    L = _lib.lib()
    n, h2, w2, c = zt.shape
    assert c == 128
    out_c = 3 if training else 2
    out = torch.empty((n, out_c, 2 * h2, 2 * w2), dtype=torch.float32, device=zt.device)
    stats = torch.empty(512, dtype=torch.float32, device=zt.device)
    ws, nb = _ws(zt.device)
    with torch.cuda.device(zt.device):
        _lib.check(L.dbb_head_tail_fwd(zt.data_ptr(), n, h2, w2, gamma.data_ptr(), beta.data_ptr(),
                                       rm.data_ptr() if rm is not None else None, rv.data_ptr() if rv is not None else None,
                                       int(training), w2b.data_ptr(), w2t.data_ptr(), b2b.data_ptr(), b2t.data_ptr(), float(k),
                                       out_c, out.data_ptr(), stats.data_ptr(), ws.data_ptr(), nb, _lib.stream_ptr()), "head_tail_fwd")
    return out, stats


def head_tail_bwd(zt, gamma, stats, w2b, w2t, out, dout, k=50.0):
    L = _lib.lib()
    n, h2, w2, c = zt.shape
    dev = zt.device
    d_zt = torch.empty_like(zt)
    f = lambda m: torch.empty(m, dtype=torch.float32, device=dev)
    dgamma, dbeta, dw2b, dw2t, db2b, db2t = f(128), f(128), f(256), f(256), f(1), f(1)
    ws, nb = _ws(dev)
    with torch.cuda.device(dev):
        _lib.check(L.dbb_head_tail_bwd(zt.data_ptr(), n, h2, w2, gamma.data_ptr(), stats.data_ptr(), w2b.data_ptr(), w2t.data_ptr(),
                                       out.data_ptr(), dout.contiguous().data_ptr(), float(k), d_zt.data_ptr(), dgamma.data_ptr(),
                                       dbeta.data_ptr(), dw2b.data_ptr(), dw2t.data_ptr(), db2b.data_ptr(), db2t.data_ptr(),
                                       ws.data_ptr(), nb, _lib.stream_ptr()), "head_tail_bwd")
    return d_zt, dgamma, dbeta, dw2b.view(64, 1, 2, 2), dw2t.view(64, 1, 2, 2), db2b, db2t
