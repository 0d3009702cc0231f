"""GPU parity of the memory-bound kernels (BatchNorm, max-pool, FPN nearest-upsample glue, fused DBHead tail) against
torch CPU fp32 on IDENTICAL (bf16-representable) inputs.  Outputs stored as bf16 carry <= 2^-9 relative rounding."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


def nhwc(t):   # (N,C,H,W) f32 cpu -> (N,H,W,C) bf16 cuda
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()


def nchw(t):   # (N,H,W,C) bf16 cuda -> (N,C,H,W) f32 cpu
    return t.float().permute(0, 3, 1, 2).contiguous().cpu()


def close_bf16(got, want, what, extra=0.0):
    got, want = got.double(), want.double()
    err = (got - want).abs()
    tol = 2.0 ** -8 * want.abs() + (2.0 ** -9 + extra) * want.abs().max() * 1e-2 + 1e-7
    bad = (err > tol)
    assert not bad.any(), (what, err.max().item(), want.abs().max().item(), int(bad.sum()))


@pytest.mark.parametrize("shape", [(2, 64, 12, 20), (3, 128, 7, 9), (2, 512, 4, 5), (1, 256, 16, 16)])
@pytest.mark.parametrize("with_res", [False, True])
def test_batchnorm_relu_train_fwd_bwd(shape, with_res):
    from db_text_minimal_b200 import _ops
    n, c, h, w = shape
    g = torch.Generator().manual_seed(c + h)
    z = bf(torch.randn(shape, generator=g) * 2 + 0.5)
    res = bf(torch.randn(shape, generator=g)) if with_res else None
    gamma = torch.rand(c, generator=g) + 0.5
    beta = torch.randn(c, generator=g) * 0.2
    rm, rv = torch.zeros(c), torch.ones(c)
    zr = z.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rr = res.clone().requires_grad_(True) if with_res else None
    rm_ref, rv_ref = rm.clone(), rv.clone()
    y = F.batch_norm(zr, rm_ref, rv_ref, gr, br, True, 0.1, 1e-5)
    if with_res:
        y = y + rr
    y = F.relu(y)
    rmd, rvd = rm.cuda(), rv.cuda()
    out, stats = _ops.bn_fwd(nhwc(z), gamma.cuda(), beta.cuda(), rmd, rvd, True, nhwc(res) if with_res else None, True)
    close_bf16(nchw(out), y.detach(), "bn fwd")
    torch.testing.assert_close(rmd.cpu(), rm_ref, rtol=1e-5, atol=1e-6)      # running stats (momentum 0.1, unbiased var)
    torch.testing.assert_close(rvd.cpu(), rv_ref, rtol=1e-5, atol=1e-6)
    # backward on the same activation mask
    dout = bf(torch.randn(shape, generator=g))
    act_ref = bf(y.detach())
    y.backward(dout)
    dz, dres, dgamma, dbeta = _ops.bn_bwd(nhwc(dout), nhwc(act_ref), nhwc(z), gamma.cuda(), stats, want_dres=with_res)
    close_bf16(nchw(dz), zr.grad, "bn dz", extra=1.0)
    torch.testing.assert_close(dgamma.cpu(), gr.grad, rtol=2e-4, atol=2e-4 * gr.grad.abs().max().item())
    torch.testing.assert_close(dbeta.cpu(), br.grad, rtol=2e-4, atol=2e-4 * br.grad.abs().max().item())
    if with_res:
        close_bf16(nchw(dres), rr.grad, "bn dres")


def test_batchnorm_eval():
    from db_text_minimal_b200 import _ops
    g = torch.Generator().manual_seed(3)
    z = bf(torch.randn((2, 64, 9, 11), generator=g))
    gamma, beta = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g)
    rm, rv = torch.randn(64, generator=g) * 0.1, torch.rand(64, generator=g) + 0.5
    y = F.batch_norm(z, rm, rv, gamma, beta, False, 0.1, 1e-5)
    out, _ = _ops.bn_fwd(nhwc(z), gamma.cuda(), beta.cuda(), rm.cuda(), rv.cuda(), False, None, False)
    close_bf16(nchw(out), y, "bn eval")


@pytest.mark.parametrize("shape", [(2, 64, 32, 40), (1, 64, 27, 35), (2, 128, 9, 9)])
def test_maxpool_fwd_bwd(shape):
    from db_text_minimal_b200 import _ops
    n, c, h, w = shape
    g = torch.Generator().manual_seed(h)
    x = bf(F.relu(torch.randn(shape, generator=g)))          # ReLU output: many exact ties at 0
    xr = x.clone().requires_grad_(True)
    y = F.max_pool2d(xr, 3, 2, 1)
    yd, am = _ops.maxpool_fwd(nhwc(x))
    assert torch.equal(nchw(yd), y.detach())                   # exact
    dy = bf(torch.randn(y.shape, generator=g))
    y.backward(dy)
    dx = _ops.maxpool_bwd(nhwc(dy), am, h, w)
    close_bf16(nchw(dx), xr.grad, "maxpool bwd")              # same arg-max convention on ties (first in scan order)


@pytest.mark.parametrize("sizes", [((20, 20), (40, 40)), ((20, 25), (39, 50)), ((7, 5), (54, 33)), ((3, 4), (5, 7)), ((9, 13), (18, 25))])
def test_nearest_upsample_add_cat_bwd(sizes):
    """F.interpolate(mode='nearest') index rule, including non-integer ratios (480x636-style inputs)."""
    from db_text_minimal_b200 import _ops
    (hs, ws), (h, w) = sizes
    g = torch.Generator().manual_seed(hs * 7 + w)
    xs = bf(torch.randn((2, 64, hs, ws), generator=g))
    y = bf(torch.randn((2, 64, h, w), generator=g))
    xr = xs.clone().requires_grad_(True)
    up = F.interpolate(xr, size=(h, w))
    out = _ops.upsample_add(nhwc(xs), nhwc(y))
    close_bf16(nchw(out), (up + y).detach(), "upsample_add")
    cat = torch.zeros((2, h, w, 256), dtype=torch.bfloat16, device="cuda")
    _ops.upsample_into(nhwc(xs), cat, 128)
    assert torch.equal(nchw(cat[..., 128:192].contiguous()), up.detach())       # pure gather: exact
    assert float(cat[..., :128].abs().sum()) == 0 and float(cat[..., 192:].abs().sum()) == 0
    dbig = bf(torch.randn((2, 256, h, w), generator=g))
    up.backward(dbig[:, 64:128])
    dxs = _ops.upsample_bwd(nhwc(dbig), 64, 64, hs, ws)
    close_bf16(nchw(dxs), xr.grad, "upsample bwd", extra=2.0)
    # accumulate flavour
    dxs2 = _ops.upsample_bwd(nhwc(dbig), 64, 64, hs, ws, d_xs=dxs.clone())
    close_bf16(nchw(dxs2), 2 * xr.grad, "upsample bwd acc", extra=4.0)


def _head_tail_ref(zt, gamma, beta, w2b, w2t, b2b, b2t, training, rm=None, rv=None, k=50.0):
    outs = []
    for br, (w2, b2) in enumerate(((w2b, b2b), (w2t, b2t))):
        sl = slice(br * 64, br * 64 + 64)
        rmb = rm[sl].clone() if rm is not None else torch.zeros(64)
        rvb = rv[sl].clone() if rv is not None else torch.ones(64)
        a = F.relu(F.batch_norm(zt[:, sl], rmb, rvb, gamma[sl], beta[sl], training, 0.1, 1e-5))
        outs.append(torch.sigmoid(F.conv_transpose2d(a, w2, b2, stride=2)))
    if training:
        outs.append(torch.reciprocal(1 + torch.exp(-k * (outs[0] - outs[1]))))
    return torch.cat(outs, 1)


@pytest.mark.parametrize("shape", [(2, 16, 24), (1, 9, 70), (3, 5, 64), (2, 40, 130), (4, 104, 192)])   # last: staged ring wraps
def test_head_tail_fwd_bwd(shape):
    """a-5 / a-6: BN+ReLU -> ConvT(64->1) x2 -> sigmoid -> step.  fp32 outputs: P, T within 1e-4 rel (north_star)."""
    from db_text_minimal_b200 import _ops
    n, h2, w2 = shape
    g = torch.Generator().manual_seed(h2 * w2)
    zt = bf(torch.randn((n, 128, h2, w2), generator=g) * 1.5)
    gamma, beta = torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g) * 0.3
    w2b, w2t = torch.randn((64, 1, 2, 2), generator=g) * 0.15, torch.randn((64, 1, 2, 2), generator=g) * 0.15
    b2b, b2t = torch.randn(1, generator=g) * 0.1, torch.randn(1, generator=g) * 0.1
    leaves = [t.clone().requires_grad_(True) for t in (zt, gamma, beta, w2b, w2t, b2b, b2t)]
    ref = _head_tail_ref(*leaves, True)
    cu = lambda t: t.cuda()
    out, stats = _ops.head_tail_fwd(nhwc(zt), cu(gamma), cu(beta), None, None, True, cu(w2b), cu(w2t), cu(b2b), cu(b2t))
    o = out.cpu()
    torch.testing.assert_close(o[:, :2], ref.detach()[:, :2], rtol=1e-4, atol=1e-6)          # P, T
    # B = step(P, T) amplifies by k=50: judge it against the step of the kernel's own P, T (SURVEY hard part 2) ...
    torch.testing.assert_close(o[:, 2], torch.reciprocal(1 + torch.exp(-50.0 * (o[:, 0] - o[:, 1]))), rtol=1e-4, atol=1e-7)
    # ... and end to end with the amplified tolerance 50 * 1e-4
    torch.testing.assert_close(o[:, 2], ref.detach()[:, 2], rtol=5e-3, atol=1e-6)
    dout = torch.randn(ref.shape, generator=g) * 1e-2
    ref.backward(dout)
    d_zt, dgamma, dbeta, dw2b, dw2t, db2b, db2t = _ops.head_tail_bwd(nhwc(zt), cu(gamma), stats, cu(w2b), cu(w2t), out, cu(dout))
    gz = leaves[0].grad
    scale = gz.abs().max().item()
    assert (nchw(d_zt) - gz).abs().max().item() <= 1.5 * 2.0 ** -8 * scale + 5e-3 * scale      # bf16 store + k=50 amplification
    for got, want, nm in ((dgamma, leaves[1].grad, "dgamma"), (dbeta, leaves[2].grad, "dbeta"), (dw2b, leaves[3].grad, "dw2b"),
                          (dw2t, leaves[4].grad, "dw2t"), (db2b, leaves[5].grad, "db2b"), (db2t, leaves[6].grad, "db2t")):
        s = want.abs().max().item() + 1e-12
        assert (got.cpu().view_as(want) - want).abs().max().item() <= 5e-3 * s, nm


def test_head_tail_eval_two_channels():
    from db_text_minimal_b200 import _ops
    g = torch.Generator().manual_seed(11)
    zt = bf(torch.randn((2, 128, 10, 12), generator=g))
    gamma, beta = torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g) * 0.3
    rm, rv = torch.randn(128, generator=g) * 0.2, torch.rand(128, generator=g) + 0.5
    w2b, w2t = torch.randn((64, 1, 2, 2), generator=g) * 0.15, torch.randn((64, 1, 2, 2), generator=g) * 0.15
    b2b, b2t = torch.randn(1, generator=g) * 0.1, torch.randn(1, generator=g) * 0.1
    ref = _head_tail_ref(zt, gamma, beta, w2b, w2t, b2b, b2t, False, rm, rv)
    cu = lambda t: t.cuda()
    out, _ = _ops.head_tail_fwd(nhwc(zt), cu(gamma), cu(beta), cu(rm), cu(rv), False, cu(w2b), cu(w2t), cu(b2b), cu(b2t))
    assert out.shape[1] == 2
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-4, atol=1e-6)


def test_step_function_matches_autograd():
    """a-6: DBHead.step_function drop-in (k B^2 e derivative)."""
    from db_text_minimal_b200.modules.segmentation_head import DBHead
    head = DBHead(256, 2)
    g = torch.Generator().manual_seed(0)
    p, t = torch.rand((2, 1, 32, 32), generator=g), torch.rand((2, 1, 32, 32), generator=g)
    pr, tr = p.clone().requires_grad_(True), t.clone().requires_grad_(True)
    ref = torch.reciprocal(1 + torch.exp(-50 * (pr - tr)))
    go = torch.randn(ref.shape, generator=g)
    ref.backward(go)
    pc, tc = p.cuda().requires_grad_(True), t.cuda().requires_grad_(True)
    b = head.step_function(pc, tc)
    b.backward(go.cuda())
    torch.testing.assert_close(b.detach().cpu(), ref.detach(), rtol=1e-4, atol=1e-30)
    scale = pr.grad.abs().max().item()
    assert (pc.grad.cpu() - pr.grad).abs().max().item() <= 1e-4 * scale
    assert (tc.grad.cpu() - tr.grad).abs().max().item() <= 1e-4 * scale
