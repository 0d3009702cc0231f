"""GPU parity: tcgen05 implicit-GEMM convolutions (through the C ABI) vs torch CPU fp32 on bf16-representable inputs.

Inputs and weights are rounded to bf16 first, so the only differences are fp32 accumulation order and the final
bf16 rounding of the stored activation: |err| <= 2^-8 |y| + accumulation noise."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (n, h, w, cin, cout, ks, stride, pad)
CONV_CASES = [
    (2, 40, 48, 64, 64, 3, 1, 1),
    (2, 40, 48, 64, 128, 3, 2, 1),
    (2, 40, 48, 128, 64, 1, 1, 0),
    (2, 40, 48, 64, 128, 1, 2, 0),
    (1, 24, 32, 256, 256, 3, 1, 1),
    (2, 18, 25, 64, 64, 3, 1, 1),       # ragged: not a multiple of any tile
    (2, 18, 25, 128, 256, 3, 2, 1),
    (3, 5, 7, 512, 64, 1, 1, 0),        # tiny maps: tiles span several images
    (16, 20, 20, 512, 512, 3, 1, 1),
    # 64 -> 64 3x3: shifted-window (halo64) fprop/dgrad and row-ring (wgrad_row64) kernels on awkward extents
    (3, 7, 33, 64, 64, 3, 1, 1),
    (1, 50, 240, 64, 64, 3, 1, 1),      # widest image the row-ring kernel takes (KP + 2 <= 256)
    (1, 20, 250, 64, 64, 3, 1, 1),      # wider: falls back to the generic split-K wgrad
    (20, 12, 16, 64, 64, 3, 1, 1),      # more images than row chunks per image
    (2, 1, 40, 64, 64, 3, 1, 1),        # single-row images
]
CONVT_CASES = [(2, 20, 24, 64, 64), (1, 9, 13, 64, 64), (2, 40, 40, 128, 64)]


def bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


def check_bf16(got, want, what):
    got, want = got.double(), want.double()
    scale = want.abs().max().item() + 1e-30
    err = (got - want).abs()
    # bf16 store: half an ulp = 2^-9 relative to the element, bounded here by the tensor max
    assert err.max().item() <= 2.0 ** -8 * scale + 1e-6, (what, err.max().item(), scale)
    assert err.mean().item() <= 2.0 ** -10 * scale, (what, "mean", err.mean().item(), scale)


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fprop_dgrad_wgrad(case):
    from db_text_minimal_b200 import _ops
    n, h, w, cin, cout, ks, st, pad = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = bf(torch.randn((n, cin, h, w), generator=g))
    wt = bf(torch.randn((cout, cin, ks, ks), generator=g) / (cin * ks * ks) ** 0.5)
    bias = torch.randn((cout,), generator=g)
    y_ref = F.conv2d(x, wt, bias, stride=st, padding=pad)
    ho, wo = y_ref.shape[2:]
    xd = _ops.to_nhwc_bf16(x.cuda())
    y = _ops.conv2d_raw(0, xd, wt.cuda(), bias.cuda(), n, h, w, cin, cout, ks, st, pad, (n, ho, wo, cout))
    check_bf16(_ops.to_nchw_f32(y).cpu(), y_ref, "fprop")
    # dgrad
    dy = bf(torch.randn((n, cout, ho, wo), generator=g))
    dx_ref = torch.nn.grad.conv2d_input(x.shape, wt, dy, stride=st, padding=pad)
    dyd = _ops.to_nhwc_bf16(dy.cuda())
    dx = _ops.conv2d_raw(1, dyd, wt.cuda(), None, n, h, w, cin, cout, ks, st, pad, (n, h, w, cin))
    check_bf16(_ops.to_nchw_f32(dx).cpu(), dx_ref, "dgrad")
    # wgrad (fp32 output)
    dw_ref = torch.nn.grad.conv2d_weight(x, wt.shape, dy, stride=st, padding=pad)
    dw = _ops.conv2d_wgrad_raw(0, xd, dyd, n, h, w, cin, cout, ks, st, pad).cpu()
    scale = dw_ref.abs().max().item()
    assert (dw - dw_ref).abs().max().item() <= 2e-4 * scale, ("wgrad", (dw - dw_ref).abs().max().item(), scale)


@pytest.mark.parametrize("case", CONVT_CASES)
def test_convtranspose_fprop_dgrad_wgrad(case):
    from db_text_minimal_b200 import _ops
    n, h, w, cin, cout = case
    g = torch.Generator().manual_seed(sum(case))
    x = bf(torch.randn((n, cin, h, w), generator=g))
    wt = bf(torch.randn((cin, cout, 2, 2), generator=g) / cin ** 0.5)
    bias = torch.randn((cout,), generator=g)
    y_ref = F.conv_transpose2d(x, wt, bias, stride=2)
    xd = _ops.to_nhwc_bf16(x.cuda())
    y = _ops.conv2d_raw(2, xd, wt.cuda(), bias.cuda(), n, h, w, cin, cout, 2, 2, 0, (n, 2 * h, 2 * w, cout))
    check_bf16(_ops.to_nchw_f32(y).cpu(), y_ref, "convT fprop")
    dy = bf(torch.randn((n, cout, 2 * h, 2 * w), generator=g))
    dx_ref = F.conv2d(dy, wt, None, stride=2)          # adjoint of conv_transpose2d
    dyd = _ops.to_nhwc_bf16(dy.cuda())
    dx = _ops.conv2d_raw(3, dyd, wt.cuda(), None, n, h, w, cin, cout, 2, 2, 0, (n, h, w, cin))
    check_bf16(_ops.to_nchw_f32(dx).cpu(), dx_ref, "convT dgrad")
    xr = x.clone().requires_grad_(False)
    wr = wt.clone().requires_grad_(True)
    F.conv_transpose2d(xr, wr, None, stride=2).backward(dy)
    dw = _ops.conv2d_wgrad_raw(2, xd, dyd, n, h, w, cin, cout, 2, 2, 0).cpu()
    scale = wr.grad.abs().max().item()
    assert (dw - wr.grad).abs().max().item() <= 2e-4 * scale


def test_layout_roundtrip():
    from db_text_minimal_b200 import _ops
    x = bf(torch.randn(2, 72, 13, 17))
    y = _ops.to_nchw_f32(_ops.to_nhwc_bf16(x.cuda())).cpu()
    assert torch.equal(x, y)
