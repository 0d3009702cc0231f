"""Runs the post-processing front on config-4-like maps a few times (target of `ncu -k regex:ccl_`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from db_text_minimal_b200 import SegDetectorRepresenter, _lib, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
s = 1024
maps = np.stack([((synth.prob_map(s, s, 100 + i) - 0.45) * 8).clip(0, 1) for i in range(8)])
P = torch.from_numpy(np.concatenate([maps] * (n // 8)))[:, None].cuda()
rep = SegDetectorRepresenter(thresh=0.25, box_thresh=0.5, unclip_ratio=1.5)
for _ in range(3):
    rep._front_device(P)
torch.cuda.synchronize()
_lib.profile_enable(True)
for _ in range(3):
    rep._front_device(P)
kern = _lib.profile_report()
_lib.profile_enable(False)
print({k["name"]: round(k["ms"] / k["launches"], 4) for k in kern}, "total/front", round(sum(k["ms"] for k in kern) / 3, 3))
