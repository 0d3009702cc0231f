"""Diagnostic run on the GPU box: every conv flavour, prints errors instead of stopping at the first failure."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from db_text_minimal_b200 import _ops, _lib


def bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


def rel(got, want):
    return ((got.double() - want.double()).abs().max() / (want.double().abs().max() + 1e-30)).item()


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    cases = [(1, 8, 16, 64, 64, 1, 1, 0), (2, 40, 48, 64, 64, 3, 1, 1), (2, 40, 48, 64, 128, 3, 2, 1),
             (2, 40, 48, 128, 64, 1, 1, 0), (2, 40, 48, 64, 128, 1, 2, 0), (1, 24, 32, 256, 256, 3, 1, 1),
             (2, 18, 25, 64, 64, 3, 1, 1), (3, 5, 7, 512, 64, 1, 1, 0), (16, 20, 20, 512, 512, 3, 1, 1)]
    for case in cases:
        n, h, w, cin, cout, ks, st, pad = case
        g = torch.Generator().manual_seed(1)
        x = bf(torch.randn((n, cin, h, w), generator=g))
        wt = bf(torch.randn((cout, cin, ks, ks), generator=g) / (cin * ks * ks) ** 0.5)
        bias = torch.randn((cout,), generator=g)
        y_ref = F.conv2d(x, wt, bias, stride=st, padding=pad)
        ho, wo = y_ref.shape[2:]
        try:
            xd = _ops.to_nhwc_bf16(x.cuda())
            y = _ops.conv2d_raw(0, xd, wt.cuda(), bias.cuda(), n, h, w, cin, cout, ks, st, pad, (n, ho, wo, cout))
            torch.cuda.synchronize()
            e0 = rel(_ops.to_nchw_f32(y).cpu(), y_ref)
            dy = bf(torch.randn((n, cout, ho, wo), generator=g))
            dyd = _ops.to_nhwc_bf16(dy.cuda())
            dx = _ops.conv2d_raw(1, dyd, wt.cuda(), None, n, h, w, cin, cout, ks, st, pad, (n, h, w, cin))
            torch.cuda.synchronize()
            e1 = rel(_ops.to_nchw_f32(dx).cpu(), torch.nn.grad.conv2d_input(x.shape, wt, dy, stride=st, padding=pad))
            dw = _ops.conv2d_wgrad_raw(0, xd, dyd, n, h, w, cin, cout, ks, st, pad)
            torch.cuda.synchronize()
            e2 = rel(dw.cpu(), torch.nn.grad.conv2d_weight(x, wt.shape, dy, stride=st, padding=pad))
            print("conv", case, "fprop %.3e dgrad %.3e wgrad %.3e" % (e0, e1, e2), flush=True)
        except Exception as e:
            print("conv", case, "EXC", repr(e), flush=True)
            traceback.print_exc()
            return
    for case in [(2, 20, 24, 64, 64), (1, 9, 13, 64, 64)]:
        n, h, w, cin, cout = case
        g = torch.Generator().manual_seed(2)
        x = bf(torch.randn((n, cin, h, w), generator=g))
        wt = bf(torch.randn((cin, cout, 2, 2), generator=g) / cin ** 0.5)
        bias = torch.randn((cout,), generator=g)
        try:
            xd = _ops.to_nhwc_bf16(x.cuda())
            y = _ops.conv2d_raw(2, xd, wt.cuda(), bias.cuda(), n, h, w, cin, cout, 2, 2, 0, (n, 2 * h, 2 * w, cout))
            e0 = rel(_ops.to_nchw_f32(y).cpu(), F.conv_transpose2d(x, wt, bias, stride=2))
            dy = bf(torch.randn((n, cout, 2 * h, 2 * w), generator=g))
            dyd = _ops.to_nhwc_bf16(dy.cuda())
            dx = _ops.conv2d_raw(3, dyd, wt.cuda(), None, n, h, w, cin, cout, 2, 2, 0, (n, h, w, cin))
            e1 = rel(_ops.to_nchw_f32(dx).cpu(), F.conv2d(dy, wt, None, stride=2))
            wr = wt.clone().requires_grad_(True)
            F.conv_transpose2d(x, wr, None, stride=2).backward(dy)
            dw = _ops.conv2d_wgrad_raw(2, xd, dyd, n, h, w, cin, cout, 2, 2, 0)
            e2 = rel(dw.cpu(), wr.grad)
            print("convT", case, "fprop %.3e dgrad %.3e wgrad %.3e" % (e0, e1, e2), flush=True)
        except Exception as e:
            print("convT", case, "EXC", repr(e), flush=True)
            traceback.print_exc()
            return
    # timing of the big layers (config 2 shapes)
    for case in [(16, 160, 160, 256, 256, 3, 1, 1), (16, 160, 160, 64, 64, 3, 1, 1), (16, 160, 160, 256, 128, 3, 1, 1),
                 (16, 80, 80, 128, 128, 3, 1, 1), (16, 40, 40, 256, 256, 3, 1, 1), (16, 20, 20, 512, 512, 3, 1, 1)]:
        n, h, w, cin, cout, ks, st, pad = case
        x = torch.randn((n, h, w, cin), device="cuda").to(torch.bfloat16)
        dy = torch.randn((n, h, w, cout), device="cuda").to(torch.bfloat16)
        wt = torch.randn((cout, cin, ks, ks), device="cuda") * 0.02
        fl = 2.0 * n * h * w * cin * cout * ks * ks
        for name, fn in (("fprop", lambda: _ops.conv2d_raw(0, x, wt, None, n, h, w, cin, cout, ks, st, pad, (n, h, w, cout))),
                         ("dgrad", lambda: _ops.conv2d_raw(1, dy, wt, None, n, h, w, cin, cout, ks, st, pad, (n, h, w, cin))),
                         ("wgrad", lambda: _ops.conv2d_wgrad_raw(0, x, dy, n, h, w, cin, cout, ks, st, pad))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print("time", case, name, "%.3f ms  %.1f TFLOP/s" % (ms, fl / ms / 1e9), flush=True)


if __name__ == "__main__":
    main()
