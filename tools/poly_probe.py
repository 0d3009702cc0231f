"""Times SegDetectorRepresenter.__call__ in polygon mode on config-4-like maps (scratch tool for gpurun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from db_text_minimal_b200 import SegDetectorRepresenter, synth

n, s = 64, 1024
maps = np.stack([((synth.prob_map(s, s, 100 + i) - 0.45) * 8).clip(0, 1) for i in range(8)])
P = torch.from_numpy(np.concatenate([maps] * (n // 8)))[:, None].cuda()
rep = SegDetectorRepresenter(thresh=0.25, box_thresh=0.5, unclip_ratio=1.5)
shape = {"shape": [(s, s)] * n}
for mode in (True, False):
    rep(shape, P, is_output_polygon=mode)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        boxes, scores = rep(shape, P, is_output_polygon=mode)
    torch.cuda.synchronize()
    print("polygon" if mode else "box", "mode: %.1f ms per batch" % ((time.perf_counter() - t0) * 1e3 / 3), "outputs per image", np.mean([len(b) if mode else int((np.asarray(b).reshape(len(b), -1) != 0).any(1).sum()) for b in boxes]))
