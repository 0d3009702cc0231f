"""Layer-by-layer comparison of the executor (bf16 / fp32 mode) with the CPU oracle on a self-conditioned network.
Scratch tool for `gpurun`:  python tools/layer_probe.py [steps]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import db_oracle as O  # noqa: E402
from db_text_minimal_b200 import DBLoss, _lib  # noqa: E402
from db_text_minimal_b200.models import DBTextModel, _aligned_workspace, _ptr_array  # noqa: E402
from db_text_minimal_b200.optim import FlatAdam  # noqa: E402


def l2rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-300)).item()


def read_taps(m, x, training, names):
    L = _lib.lib()
    n, _, h, w = x.shape
    plan = m._plan(n, h, w, training)
    raw, ws = _aligned_workspace(plan.ws_bytes, x.device)
    out = torch.empty((n, plan.out_c, h, w), device=x.device)
    pa = _ptr_array([p.data_ptr() for p in m._param_list()])
    ba = _ptr_array([b.data_ptr() for b in m._buffer_list()])
    _lib.check(L.dbb_net_forward(plan.handle, x.data_ptr(), pa, ba, out.data_ptr(), ws, plan.ws_bytes, _lib.stream_ptr()), "fwd")
    res = {}
    for nm in names:
        shp = (C.c_int64 * 4)()
        if L.dbb_net_debug_shape(plan.handle, nm.encode(), shp):
            continue
        t = torch.empty(tuple(shp), device=x.device)
        _lib.check(L.dbb_net_debug_read(plan.handle, nm.encode(), ws, t.data_ptr(), _lib.stream_ptr()), nm)
        res[nm] = t.cpu()
    torch.cuda.synchronize()
    return out.cpu(), res


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    params = O.init_params(O.COND_SEED)
    m = DBTextModel(precision="bf16", pretrained=False)
    m.load_state_dict(params)
    m = m.cuda().train()
    opt = FlatAdam(m, lr=0.005)
    crit = DBLoss(reduction="mean")
    for it in range(steps):
        x, g = O.synth_text_batch(4, 128, 128, it)
        ls = crit(m(x.cuda()), torch.from_numpy(g).cuda())
        opt.zero_grad(); ls[-1].backward(); opt.step()
    print("trained; loss", float(ls[-1]))
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    m32 = DBTextModel(precision="fp32", pretrained=False)
    m32.load_state_dict(sd)
    m32 = m32.cuda()
    names = ["x1"] + [f"block{i}.out" for i in range(8)] + ["p5", "l4", "p4", "l3", "p3", "l2", "cat", "af", "ah"]
    for training in (False, True):
        x, _ = O.synth_text_batch(1, 640, 640, 901)
        taps, taps_q = {}, {}
        with torch.no_grad():
            ref = O.dbnet_forward(sd, x, training, taps=taps)
            refq = O.dbnet_forward(sd, x, training, quant=O.bf16_round, taps=taps_q)
        print("==== training" if training else "==== eval", "| oracle bf16-emulation vs oracle fp32:", [round(l2rel(refq[:, c], ref[:, c]), 5) for c in range(2)])
        for tag, model in (("bf16", m), ("fp32", m32)):
            model.train(training)
            sdc = {k: v.clone() for k, v in model.state_dict().items()}
            out, got = read_taps(model, x.cuda(), training, names)
            model.load_state_dict(sdc)      # undo the running-statistics update of a training-mode probe
            print(tag, "P,T", [l2rel(out[:, c], ref[:, c]) for c in range(2)])
            for nm in names:
                if nm in got and nm in taps:
                    print(f"   {nm:12s} {l2rel(got[nm], taps[nm]):.3e}   (oracle bf16 emulation {l2rel(taps_q[nm], taps[nm]):.3e})  absmax {taps[nm].abs().max():.3g}")


if __name__ == "__main__":
    main()
