"""ConvTranspose2d(64, 64, 2, 2) at the head's shape (16 x 160 x 160) and the stem through the single-operator C ABI
(target of `ncu -k regex:igemm_persist`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from db_text_minimal_b200 import _ops, _lib

n, h, w = 16, 160, 160
x = (torch.randn(n, h, w, 64, device="cuda")).to(torch.bfloat16)
wt = torch.randn(64, 64, 2, 2, device="cuda") * 0.1
b = torch.randn(64, device="cuda")


def run():
    return _ops.conv2d_raw(2, x, wt, b, n, h, w, 64, 64, 2, 2, 0, (n, 2 * h, 2 * w, 64))


x1 = (torch.randn(n, h, w, 256, device="cuda")).to(torch.bfloat16)
w1 = torch.randn(64, 256, 1, 1, device="cuda") * 0.1


def run1():
    return _ops.conv2d_raw(0, x1, w1, b, n, h, w, 256, 64, 1, 1, 0, (n, h, w, 64))


for _ in range(3):
    run(); run1()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("convT 64->64 k2s2 16x160x160", run), ("1x1 256->64 16x160x160", run1)):
    _lib.profile_enable(True)
    for _ in range(5):
        fn()
    k = _lib.profile_report()
    _lib.profile_enable(False)
    print(name, {r["name"]: round(r["ms"] / r["launches"], 4) for r in k})
