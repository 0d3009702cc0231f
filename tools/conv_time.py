"""Times single conv launches (fprop / dgrad / wgrad) at the config-2 layer shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from db_text_minimal_b200 import _ops
cases = [(16, 160, 160, 256, 256, 3, 1, 1), (16, 160, 160, 64, 64, 3, 1, 1), (16, 160, 160, 256, 128, 3, 1, 1),
         (16, 80, 80, 128, 128, 3, 1, 1), (16, 40, 40, 256, 256, 3, 1, 1), (16, 20, 20, 512, 512, 3, 1, 1), (16, 160, 160, 64, 64, 1, 1, 0)]
import torch.nn.functional as F
# ConvTranspose2d(64,64,2,2) forward at the head resolution + the 1x1 lateral
for (n, h, w, cin, cout) in [(16, 160, 160, 64, 64)]:
    x = torch.randn((n, h, w, cin), device="cuda").to(torch.bfloat16)
    wt = torch.randn((cin, cout, 2, 2), device="cuda") * 0.02
    fn = lambda: _ops.conv2d_raw(2, x, wt, None, n, h, w, cin, cout, 2, 2, 0, (n, 2 * h, 2 * w, cout))
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record(); torch.cuda.synchronize()
    print("time convT", (n, h, w, cin, cout), "%.3f ms" % (e0.elapsed_time(e1) / 20), flush=True)
for case in cases:
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
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("time", case, name, "%.3f ms  %.1f TFLOP/s" % (ms, fl / ms / 1e9), flush=True)
