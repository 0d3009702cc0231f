"""Per-operator autograd functions over the single-operator C ABI (include/dbb200.h) -- the stand-alone execution path of
the sub-modules (ConvBnRelu, BasicBlock, ResNet, FPN, DBHead), i.e. what a user gets when calling a piece of the network
outside DBTextModel.  Inside this layer activations are (N, H, W, C) bfloat16 CUDA tensors; modules convert at their
NCHW float32 boundary.  DBTextModel.forward does NOT go through here: it runs the fused executor (csrc/net.cu)."""
import torch

from . import _lib, _ops


class ToNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _ops.to_nhwc_bf16(x)

    @staticmethod
    def backward(ctx, g):
        return _ops.to_nchw_f32(g.contiguous())


class ToNCHW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _ops.to_nchw_f32(x)

    @staticmethod
    def backward(ctx, g):
        return _ops.to_nhwc_bf16(g)


class Conv(torch.autograd.Function):
    """nn.Conv2d (1x1 / 3x3, stride 1 / 2, channels multiples of 64) on NHWC bf16."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad):
        n, h, w, cin = x.shape
        cout, _, ks, _ = weight.shape
        ho, wo = (h + 2 * pad - ks) // stride + 1, (w + 2 * pad - ks) // stride + 1
        y = _ops.conv2d_raw(0, x, weight, bias, n, h, w, cin, cout, ks, stride, pad, (n, ho, wo, cout))
        ctx.save_for_backward(x, weight)
        ctx.geom = (n, h, w, cin, cout, ks, stride, pad, bias is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        n, h, w, cin, cout, ks, stride, pad, has_bias = ctx.geom
        dy = dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _ops.conv2d_raw(1, dy, weight, None, n, h, w, cin, cout, ks, stride, pad, (n, h, w, cin))
        if ctx.needs_input_grad[1]:
            dw = _ops.conv2d_wgrad_raw(0, x, dy, n, h, w, cin, cout, ks, stride, pad)
        if has_bias and ctx.needs_input_grad[2]:
            db = dy.float().sum(dim=(0, 1, 2))
        return dx, dw, db, None, None


class ConvT(torch.autograd.Function):
    """nn.ConvTranspose2d(cin, cout, 2, 2) on NHWC bf16 (channels multiples of 64)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        n, h, w, cin = x.shape
        cout = weight.shape[1]
        y = _ops.conv2d_raw(2, x, weight, bias, n, h, w, cin, cout, 2, 2, 0, (n, 2 * h, 2 * w, cout))
        ctx.save_for_backward(x, weight)
        ctx.geom = (n, h, w, cin, cout, bias is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        n, h, w, cin, cout, has_bias = ctx.geom
        dy = dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _ops.conv2d_raw(3, dy, weight, None, n, h, w, cin, cout, 2, 2, 0, (n, h, w, cin))
        if ctx.needs_input_grad[1]:
            dw = _ops.conv2d_wgrad_raw(2, x, dy, n, h, w, cin, cout, 2, 2, 0)
        if has_bias and ctx.needs_input_grad[2]:
            db = dy.float().sum(dim=(0, 1, 2))
        return dx, dw, db


class Stem(torch.autograd.Function):
    """ResNet.conv1: Conv2d(3, 64, 7, stride 2, pad 3, bias=False) from the NCHW float32 image to NHWC bf16."""

    @staticmethod
    def forward(ctx, img, weight):
        L = _lib.lib()
        _lib.require_cuda(img, weight)
        img = img.float().contiguous()
        n, _, h, w = img.shape
        y = torch.empty((n, (h + 1) // 2, (w + 1) // 2, 64), dtype=torch.bfloat16, device=img.device)
        nb = L.dbb_conv1_workspace(n, h, w, 0)
        raw = torch.empty(nb + 1024, dtype=torch.uint8, device=img.device)
        ptr = (raw.data_ptr() + 1023) // 1024 * 1024
        wf = weight.float().contiguous()
        with torch.cuda.device(img.device):
            _lib.check(L.dbb_conv1_fwd(img.data_ptr(), wf.data_ptr(), y.data_ptr(), n, h, w, ptr, nb, _lib.stream_ptr()), "dbb_conv1_fwd")
        ctx.save_for_backward(img)
        return y

    @staticmethod
    def backward(ctx, dy):
        (img,) = ctx.saved_tensors
        L = _lib.lib()
        n, _, h, w = img.shape
        dy = dy.contiguous()
        dw = torch.empty((64, 3, 7, 7), dtype=torch.float32, device=img.device)
        nb = L.dbb_conv1_workspace(n, h, w, 1)
        raw = torch.empty(nb + 1024, dtype=torch.uint8, device=img.device)
        ptr = (raw.data_ptr() + 1023) // 1024 * 1024
        with torch.cuda.device(img.device):
            _lib.check(L.dbb_conv1_wgrad(img.data_ptr(), dy.data_ptr(), dw.data_ptr(), n, h, w, ptr, nb, _lib.stream_ptr()), "dbb_conv1_wgrad")
        return None, dw


class BatchNorm(torch.autograd.Function):
    """nn.BatchNorm2d (+ residual) (+ ReLU) on NHWC bf16; running statistics are updated in place in training mode."""

    @staticmethod
    def forward(ctx, z, gamma, beta, running_mean, running_var, training, residual, relu):
        out, stats = _ops.bn_fwd(z, gamma.float(), beta.float(), running_mean, running_var, training, residual, relu)
        ctx.save_for_backward(z, gamma, stats, out if relu else None)
        ctx.training, ctx.has_res, ctx.relu = bool(training), residual is not None, bool(relu)
        return out

    @staticmethod
    def backward(ctx, dout):
        z, gamma, stats, act = ctx.saved_tensors
        dout = dout.contiguous()
        c = z.shape[-1]
        if ctx.training:
            dz, dres, dgamma, dbeta = _ops.bn_bwd(dout, act, z, gamma.float(), stats, want_dres=ctx.has_res)
        else:       # eval mode: fixed statistics, y = z*scale + shift  (not on any hot path: plain tensor arithmetic)
            dy = dout.float()
            if act is not None:
                dy = dy * (act > 0)
            scale, mean, inv = stats[:c], stats[2 * c:3 * c], stats[3 * c:]
            dz = (dy * scale).to(torch.bfloat16)
            dres = dy.to(torch.bfloat16) if ctx.has_res else None
            dbeta = dy.sum(dim=(0, 1, 2))
            dgamma = (dy * ((z.float() - mean) * inv)).sum(dim=(0, 1, 2))
        return dz, dgamma, dbeta, None, None, None, dres, None


class MaxPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y, am = _ops.maxpool_fwd(x)
        ctx.save_for_backward(am)
        ctx.hw = (x.shape[1], x.shape[2])
        return y

    @staticmethod
    def backward(ctx, dy):
        (am,) = ctx.saved_tensors
        return _ops.maxpool_bwd(dy.contiguous(), am, *ctx.hw)


class UpsampleAdd(torch.autograd.Function):
    """FPN._upsample_add (segmentation_body.py:79-80): nearest(xs -> y's size) + y."""

    @staticmethod
    def forward(ctx, xs, y):
        ctx.small = (xs.shape[1], xs.shape[2], xs.shape[3])
        return _ops.upsample_add(xs, y)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        hs, ws, c = ctx.small
        return _ops.upsample_bwd(g, 0, c, hs, ws), g


class UpsampleCat(torch.autograd.Function):
    """FPN._upsample_cat (segmentation_body.py:82-87): cat(p2, up(p3), up(p4), up(p5)) along channels."""

    @staticmethod
    def forward(ctx, p2, p3, p4, p5):
        n, h, w, c = p2.shape
        cat = torch.empty((n, h, w, 4 * c), dtype=torch.bfloat16, device=p2.device)
        cat[..., :c] = p2
        for i, p in enumerate((p3, p4, p5)):
            _ops.upsample_into(p.contiguous(), cat, (i + 1) * c)
        ctx.shapes = [(p.shape[1], p.shape[2]) for p in (p3, p4, p5)]
        ctx.c = c
        return cat

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        c = ctx.c
        grads = [g[..., :c].contiguous()]
        for i, (hs, ws) in enumerate(ctx.shapes):
            grads.append(_ops.upsample_bwd(g, (i + 1) * c, c, hs, ws))
        return tuple(grads)


class HeadTail(torch.autograd.Function):
    """BatchNorm + ReLU + ConvTranspose2d(64, 1, 2, 2) + Sigmoid of both branches + step function (segmentation_head.py:
    28-29, 39-44, 72-76, 106-108) on the 128-channel tensor [binarize | thresh]."""

    @staticmethod
    def forward(ctx, zt, gamma, beta, rm, rv, training, w2b, w2t, b2b, b2t, k):
        out, stats = _ops.head_tail_fwd(zt, gamma, beta, rm, rv, training, w2b.float().contiguous(), w2t.float().contiguous(),
                                        b2b.float().contiguous(), b2t.float().contiguous(), k)
        ctx.save_for_backward(zt, gamma, stats, w2b, w2t, out)
        ctx.k, ctx.training = k, bool(training)
        return out

    @staticmethod
    def backward(ctx, dout):
        if not ctx.training:
            raise _lib.DbbError("DBHead backward is only available in training mode")
        zt, gamma, stats, w2b, w2t, out = ctx.saved_tensors
        d_zt, dgamma, dbeta, dw2b, dw2t, db2b, db2t = _ops.head_tail_bwd(zt, gamma, stats, w2b.float().contiguous(),
                                                                         w2t.float().contiguous(), out, dout.float().contiguous(), ctx.k)
        return d_zt, dgamma, dbeta, None, None, None, dw2b, dw2t, db2b, db2t, None


# ------------------------------------------------------------------ helpers used by the module forwards
def conv_bn(x, conv, bn, relu=True, residual=None):
    """conv (nn.Conv2d) -> bn (nn.BatchNorm2d) [-> + residual] [-> ReLU] on NHWC bf16, with nn.BatchNorm2d's bookkeeping."""
    z = Conv.apply(x, conv.weight, conv.bias, conv.stride[0], conv.padding[0])
    return batch_norm(z, bn, relu, residual)


def batch_norm(z, bn, relu=True, residual=None):
    training = bn.training
    out = BatchNorm.apply(z, bn.weight, bn.bias, bn.running_mean, bn.running_var, training, residual, relu)
    if training and bn.num_batches_tracked is not None:
        with torch.no_grad():
            bn.num_batches_tracked += 1
    return out
