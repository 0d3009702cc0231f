"""Kernel timeline of graph-replayed training steps via torch.profiler (CUPTI): per-stream busy time, how much of the
weight-gradient stream runs concurrently with which kind of main-stream kernel.  Scratch tool for gpurun."""
import os, sys, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from db_text_minimal_b200 import DBLoss, DBTextModel, synth
from db_text_minimal_b200.graph import GraphedTrainStep
from db_text_minimal_b200.optim import FlatAdam

N, S = 16, 640
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = DBTextModel(pretrained=False).to(dev).train()
crit = DBLoss(alpha=1.0, beta=10.0, reduction="none", negative_ratio=3)
opt = FlatAdam(model, lr=0.005)
img = synth.images(N, S, S, seed=0).to(dev)
gts = torch.from_numpy(synth.gt_maps(N, S, S, seed=0)).to(dev)
step = GraphedTrainStep(model, crit, opt, tuple(img.shape), tuple(gts.shape), dev).capture(img, gts)
for _ in range(5):
    step.replay()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step.replay()
    torch.cuda.synchronize()
prof.export_chrome_trace("gpurun_out/timeline.json")
ev = [e for e in json.load(open("gpurun_out/timeline.json"))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
streams = collections.Counter(e["args"]["stream"] for e in ev)
print("kernels", len(ev), "streams", dict(streams))
# take the middle replay: split by the first kernel name
first = ev[0]["name"]
starts = [i for i, e in enumerate(ev) if e["name"] == first and e["args"]["stream"] == ev[0]["args"]["stream"]]
print("first kernel", first[:50], "occurrences", len(starts))
seg = ev[starts[1]:starts[2]] if len(starts) >= 3 else ev
t0 = seg[0]["ts"]; t1 = max(e["ts"] + e["dur"] for e in seg)
print("step span %.1f us" % (t1 - t0))
main = max(streams, key=streams.get)
rows = []
for e in seg:
    rows.append((e["ts"] - t0, e["dur"], e["args"]["stream"] == main, e["name"]))
json.dump(rows, open("gpurun_out/timeline_rows.json", "w"))
busy_main = sum(d for _, d, m, _ in rows if m); busy_side = sum(d for _, d, m, _ in rows if not m)
print("sum of durations: main %.1f us, side %.1f us" % (busy_main, busy_side))
