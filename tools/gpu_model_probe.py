"""Diagnostic: full DBTextModel forward/backward on the GPU vs the CPU oracle, layer by layer."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
from oracle import db_oracle as O
from db_text_minimal_b200 import _lib
from db_text_minimal_b200.models import DBTextModel, _DBNetFn
from db_text_minimal_b200.losses import DBLoss


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item(), ((a - b).norm() / (b.norm() + 1e-30)).item()


def debug_read(model, plan, ws_ptr, name):
    L = _lib.lib()
    shp = (C.c_int64 * 4)()
    _lib.check(L.dbb_net_debug_shape(plan.handle, name.encode(), shp), name)
    out = torch.empty(tuple(shp), dtype=torch.float32, device="cuda")
    _lib.check(L.dbb_net_debug_read(plan.handle, name.encode(), ws_ptr, out.data_ptr(), _lib.stream_ptr()), name)
    torch.cuda.synchronize()
    return out.cpu()


def oracle_intermediates(params, x, training, quant):
    c = O._Ctx(params, training, quant)
    inter = {}
    z0 = c.conv(x, "backbone.conv1", 2, 3); inter["z0"] = z0
    a0 = F.relu(c.bn(z0, "backbone.bn1")); inter["a0"] = a0
    xx = F.max_pool2d(a0, 3, 2, 1); inter["x1"] = xx
    feats = []
    i = 0
    for li in range(1, 5):
        for bi in range(2):
            xx = O._basic_block(c, xx, f"backbone.layer{li}.{bi}", 2 if (li > 1 and bi == 0) else 1)
            inter[f"block{i}.out"] = xx
            i += 1
        feats.append(xx)
    f = O.fpn_forward(c, tuple(feats)); inter["af"] = f
    # head, step by step (segmentation_head.py:35-45)
    zh, ah, zt = [], [], []
    outs = []
    for br in ("binarize", "thresh"):
        pre = "segmentation_head." + br
        z = c.conv(f, pre + ".0", 1, 1); zh.append(z)
        a = F.relu(c.bn(z, pre + ".1")); ah.append(a)
        t = c.q(c.convT(a, pre + ".3")); zt.append(t)
        yb = F.relu(c.bn(t, pre + ".4"))
        outs.append(torch.sigmoid(F.conv_transpose2d(yb, c.p[pre + ".6.weight"], c.p[pre + ".6.bias"], stride=2)))
    for nm, lst in (("zh", zh), ("ah", ah), ("zt", zt)):
        inter[nm + "0"], inter[nm + "1"] = lst
    if training:
        outs.append(O.step_function(outs[0], outs[1]))
    y = torch.cat(outs, 1)
    return y, inter


def main():
    torch.manual_seed(0)
    n, h, w = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (2, 128, 160)))
    seed = 0
    params = O.init_params(seed)
    x = O.synth_images(n, h, w, seed)
    gts = O.synth_gt_maps(n, h, w, seed)
    model = DBTextModel()
    model.load_state_dict(params, strict=True)
    model.cuda()
    for training in (False, True):
        model.train(training)
        # keep the workspace alive to read intermediates
        plan = model._plan(n, h, w, training)
        y = model(x.cuda())
        torch.cuda.synchronize()
        with torch.no_grad():
            y32, _ = oracle_intermediates(params, x, training, None)
            ybf, inter = oracle_intermediates(params, x, training, O.bf16_round)
        print(f"== training={training} out shape {tuple(y.shape)}")
        for ch, nm in enumerate(["P", "T", "B"][:y.shape[1]]):
            print(f"  {nm}: vs fp32 oracle max/l2 rel = %.3e %.3e | vs bf16-emulating oracle = %.3e %.3e" %
                  (*rel(y[:, ch].cpu(), y32[:, ch]), *rel(y[:, ch].cpu(), ybf[:, ch])))
        if training:
            ws_ptr = y.grad_fn.ws_ptr
        else:
            ws_ptr = list(model._eval_ws.values())[0][1]
        for name in ["z0", "a0", "x1"] + [f"block{i}.out" for i in range(8)] + ["af"]:
            t = debug_read(model, plan, ws_ptr, name)
            print(f"  {name:12s} vs bf16-oracle max/l2 rel = %.3e %.3e" % rel(t, inter[name]))
    # ---- backward
    model.train(True)
    model.zero_grad()
    y = model(x.cuda())
    plan = model._plan(n, h, w, True)
    ws_ptr = y.grad_fn.ws_ptr
    ws_keep = y.grad_fn.ws_raw      # keep the workspace alive past backward()
    USE_LOSS = os.environ.get("PROBE_LOSS", "0") == "1"
    g = torch.Generator().manual_seed(5)
    dout_fixed = torch.randn(y.shape, generator=g) * 1e-3
    dout_fixed[:, 2] = 0      # the step function amplifies forward differences by k=50: keep B out of this check
    if USE_LOSS:
        crit = DBLoss(reduction="mean")
        ls = crit(y, torch.from_numpy(gts).cuda())
        ls[-1].backward()
        print("losses gpu", [float(v) for v in ls])
    else:
        y.backward(dout_fixed.cuda())
    torch.cuda.synchronize()
    po = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in params.items()}
    yo, inter = oracle_intermediates(po, x, True, O.bf16_round)
    for t in inter.values():
        t.retain_grad()
    if USE_LOSS:
        res = O.db_loss(yo.detach().numpy(), gts, reduction="mean")
        print("losses oracle(bf16-emulated fwd)", res["losses"])
        yo.backward(torch.from_numpy(res["grad"]).float())
    else:
        yo.backward(dout_fixed)
    # NOTE: the workspace was released by backward(); it is still intact until the next allocation
    pairs = [("d_zt", "zt"), ("d_ah", "ah"), ("d_zh", "zh"), ("d_af", "af")] + [(f"block{i}.d_out", f"block{i}.out") for i in range(7, -1, -1)] + \
            [("d_x1", "x1"), ("d_a0", "a0"), ("d_z0", "z0")]
    for mine, theirs in pairs:
        t = debug_read(model, plan, ws_ptr, mine)
        og = torch.cat([inter[theirs + "0"].grad, inter[theirs + "1"].grad], 1) if theirs in ("zt", "ah", "zh") else inter[theirs].grad
        print(f"  {mine:14s} vs oracle grad max/l2 rel = %.3e %.3e" % rel(t, og))
    named = dict(model.named_parameters())
    worst = []
    for k, p in named.items():
        if po[k].grad is None:
            assert p.grad is None, k
            continue
        g = p.grad.cpu()
        m, l2 = rel(g, po[k].grad)
        worst.append((l2, m, k))
    worst.sort(reverse=True)
    for l2, m, k in worst[:25]:
        print("  grad %-55s l2rel %.3e maxrel %.3e" % (k, l2, m))
    print("  median l2rel %.3e" % np.median([w[0] for w in worst]))


if __name__ == "__main__":
    main()
