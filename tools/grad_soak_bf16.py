"""One-off: gradients of the bf16 product path against the fp32 mode of the same executor at random image sizes (a missing
tile row / column in a tcgen05 dgrad or wgrad kernel would show as a collapsed cosine on that layer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import test_model_fp32_gpu as T
from oracle import db_oracle as O
from db_text_minimal_b200 import DBLoss
rng = np.random.RandomState(21)
params = O.cond_params(T.GOLD)
models = {p: T.build(params, p) for p in ("fp32", "bf16")}
for ci in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    n = int(rng.randint(1, 4))
    h, w = [(int(rng.randint(3, 12)) * 32, int(rng.randint(3, 12)) * 32), (int(rng.randint(20, 90)) * 4, int(rng.randint(20, 90)) * 4),
            (int(rng.randint(70, 300)), int(rng.randint(70, 300)))][ci % 3]
    x, gts = O.synth_text_batch(n, h, w, 900 + ci)
    g = {}
    for prec, m in models.items():
        m.load_state_dict(params); m.train(); m.zero_grad(set_to_none=True)
        ls = DBLoss(alpha=1.0, beta=10.0, reduction="none", negative_ratio=3)(m(x.cuda()), torch.from_numpy(gts).cuda())
        ls[-1].backward()
        g[prec] = {k: p.grad.detach().double().cpu().flatten() for k, p in m.named_parameters() if p.grad is not None}
    keys = [k for k in g["bf16"] if k not in set(T.zero_grad_keys(list(g["bf16"]))) and g["fp32"][k].norm() > 0]
    cos = sorted((torch.nn.functional.cosine_similarity(g["bf16"][k], g["fp32"][k], dim=0).item(), k) for k in keys)
    print((n, h, w), "cosine bf16 vs fp32 mode: min %.3f (%s) p10 %.3f median %.3f" % (cos[0][0], cos[0][1], cos[len(cos) // 10][0], cos[len(cos) // 2][0]), flush=True)
