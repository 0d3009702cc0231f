"""Times the pieces of the box-mode post-processing path on config-4-like maps (scratch tool for gpurun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from db_text_minimal_b200 import SegDetectorRepresenter, _lib, synth

n, s = 64, 1024
maps = np.stack([((synth.prob_map(s, s, 100 + i) - 0.45) * 8).clip(0, 1) for i in range(8)])
P = torch.from_numpy(np.concatenate([maps] * (n // 8)))[:, None].cuda()
rep = SegDetectorRepresenter(thresh=0.25, box_thresh=0.5, unclip_ratio=1.5)
dest = [(s, s)] * n
for _ in range(2):
    rep.boxes_batch(P, dest)
torch.cuda.synchronize()
L = _lib.lib()


def t():
    torch.cuda.synchronize()
    return time.perf_counter()


t0 = t(); f = rep._front_device(P); t1 = t()
cap = max(8192, (s * s) // 16)
npts_dev = torch.empty(n, dtype=torch.int32, device="cuda")
pts_dev = torch.empty((n, cap, 2), dtype=torch.int32, device="cuda")
t2 = t()
_lib.check(L.dbb_ccl_border_points(f["ws"].data_ptr(), f["wsb"], n, s, s, pts_dev.data_ptr(), npts_dev.data_ptr(), cap, _lib.stream_ptr()), "pts")
t3 = t()
counts = torch.stack([f["ncand"], npts_dev]).cpu().numpy()
t4 = t()
nc, npts = counts
k = int(min(nc.max(), 1000)); pmax = int(npts.max()); assert pmax <= cap, (pmax, cap)
cands = np.ascontiguousarray(f["cands"][:, :k].cpu().numpy())
pts = np.ascontiguousarray(pts_dev[:, :pmax].cpu().numpy())
t5 = t()
boxes = np.empty((n, k, 4, 2), np.int16); scores = np.empty((n, k), np.float32)
dst = np.ascontiguousarray(np.asarray(dest, np.int32))
nc32, np32 = np.ascontiguousarray(nc.astype(np.int32)), np.ascontiguousarray(npts.astype(np.int32))
for threads in (1, 4, 16):
    ta = time.perf_counter()
    _lib.check(L.dbb_boxes_from_border_points(cands.ctypes.data, nc32.ctypes.data, pts.ctypes.data, np32.ctypes.data, n, k, pmax, s, s, dst.ctypes.data,
                                              1.5, 3, boxes.ctypes.data, scores.ctypes.data, None, None, threads), "boxes")
    print("host back half, threads", threads, (time.perf_counter() - ta) * 1e3, "ms")
print("front %.2f ms | alloc %.2f | points kernel %.2f | counts d2h %.2f | cands+points d2h %.2f (k=%d, pmax=%d, %.1f MB)" % (
    (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t5 - t4) * 1e3, k, pmax, (cands.nbytes + pts.nbytes) / 1e6))
print("kept boxes per image", float((boxes.reshape(n, k, -1) != 0).any(-1).sum(1).mean()), "points per image", float(npts.mean()))
ta = t(); rep.boxes_batch(P, dest); print("boxes_batch total %.2f ms" % ((t() - ta) * 1e3))
ta = t(); rep({"shape": dest}, P); print("__call__ total %.2f ms" % ((t() - ta) * 1e3))
