"""One-off soak of the executor at random image sizes (tile edges of every convolution, non-integer FPN ratios, bilinear
resize when H or W is not a multiple of 4): forward in train and eval mode against the CPU oracle on the conditioned
fixture weights, fp32 mode at 1e-4 and the bf16 product path at 3e-2.  python tools/model_soak.py [cases]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import test_model_fp32_gpu as T
from oracle import db_oracle as O

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rng = np.random.RandomState(9)
params = O.cond_params(T.GOLD)
models = {p: T.build(params, p) for p in ("fp32", "bf16")}
bad = 0
worst = {"fp32": 0.0, "bf16": 0.0}
for ci in range(cases):
    n = int(rng.randint(1, 4))
    if ci % 3 == 0:   h, w = int(rng.randint(10, 60)) * 4, int(rng.randint(10, 60)) * 4          # multiples of 4
    elif ci % 3 == 1: h, w = int(rng.randint(2, 8)) * 32, int(rng.randint(2, 8)) * 32            # multiples of 32
    else:             h, w = int(rng.randint(40, 230)), int(rng.randint(40, 230))                # anything (bilinear resize at the end)
    x, _ = O.synth_text_batch(n, h, w, 500 + ci)
    for training in (True, False):
        for prec, m in models.items():
            m.train(training)
            sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
            with torch.no_grad():
                y = m(x.cuda()).cpu()
                ref = O.dbnet_forward(sd, x, training)
            m.load_state_dict(sd)
            for ch in range(2):
                e = T.l2rel(y[:, ch], ref[:, ch])
                worst[prec] = max(worst[prec], e)
                if e > (1e-4 if prec == "fp32" else 3e-2):
                    bad += 1
                    print("FAIL", (n, h, w), "train" if training else "eval", prec, "PT"[ch], e, flush=True)
print("MODEL SOAK cases", cases, "failed", bad, "worst l2 error", worst)

# ---- gradients in the fp32 mode at a few random sizes: every parameter tensor against the float64 evaluation of the oracle
# (criterion of tests/test_model_fp32_gpu.py: e_mine <= max(1e-3, 3 e_ref), or direct agreement with the float32 oracle to 1e-3)
from db_text_minimal_b200 import DBLoss
gbad = 0
m = models["fp32"]
for ci in range(int(sys.argv[2]) if len(sys.argv) > 2 else 4):
    n = int(rng.randint(1, 3))
    h, w = (int(rng.randint(40, 180)), int(rng.randint(40, 180))) if ci % 2 else (int(rng.randint(10, 45)) * 4, int(rng.randint(10, 45)) * 4)
    x, gts = O.synth_text_batch(n, h, w, 700 + ci)
    red = "none" if ci % 2 else "mean"
    m.load_state_dict(params); m.train(); m.zero_grad(set_to_none=True)
    ls = DBLoss(alpha=1.0, beta=10.0, reduction=red, negative_ratio=3)(m(x.cuda()), torch.from_numpy(gts).cuda())
    ls[-1].backward()
    _, _, og = T.oracle_grads(params, x, gts, red)
    _, _, g64 = T.oracle_grads(params, x, gts, red, torch.float64)
    mine = {k: p.grad.detach().cpu() for k, p in m.named_parameters() if p.grad is not None}
    keys = [k for k in mine if k in og and k not in set(T.zero_grad_keys(list(mine)))]
    em = [T.l2rel(mine[k], g64[k]) for k in keys]; er = [T.l2rel(og[k], g64[k]) for k in keys]; ed = [T.l2rel(mine[k], og[k]) for k in keys]
    fails = [(k, a, b, c) for k, a, b, c in zip(keys, em, er, ed) if a > max(1e-3, 3 * b) and c > 1e-3]
    gbad += len(fails)
    print("grad case", (n, h, w), red, "ours median %.2e max %.2e | oracle-fp32 median %.2e max %.2e | failing tensors %d" %
          (np.median(em), max(em), np.median(er), max(er), len(fails)), fails[:2], flush=True)
print("GRAD SOAK failing tensors", gbad)
