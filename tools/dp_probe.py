"""2-rank smoke of the data-parallel path with stage prints (debugging aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
t0 = time.time()
def log(*a):
    print(f"[r{os.environ.get('RANK')} {time.time()-t0:6.1f}s]", *a, flush=True)
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
log("init pg")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
log("pg ready")
x = torch.ones(4, device="cuda") * (dist.get_rank() + 1)
dist.all_reduce(x); torch.cuda.synchronize(); log("allreduce ok", x.tolist())
from db_text_minimal_b200 import DBLoss, DBTextModel, synth
from db_text_minimal_b200.dist import GradSync
torch.manual_seed(0)
model = DBTextModel(pretrained=False).cuda().train()
sync = GradSync(model)
crit = DBLoss(reduction="none")
opt = torch.optim.Adam(model.parameters(), lr=0.005, fused=True)
N, S = 4, 256
img = synth.images(N, S, S, seed=dist.get_rank()).cuda()
gts = torch.from_numpy(synth.gt_maps(N, S, S, seed=dist.get_rank())).cuda()
for it in range(3):
    opt.zero_grad(set_to_none=True)
    log("fwd", it)
    y = model(img)
    l = crit(y, gts)[-1]
    log("bwd", it)
    l.backward()
    torch.cuda.synchronize(); log("bwd done", it, float(l))
    opt.step()
    torch.cuda.synchronize(); log("step done", it)
# gradients identical across ranks after sync
g = torch.cat([p.grad.flatten() for p in model.parameters() if p.grad is not None])
ref = g.clone(); dist.broadcast(ref, 0)
log("grad equal across ranks:", bool(torch.equal(ref, g)), "weights equal:",
    bool(torch.equal(*[t for t in [torch.cat([p.detach().flatten() for p in model.parameters()])] * 2])))
w = torch.cat([p.detach().flatten() for p in model.parameters()]); wr = w.clone(); dist.broadcast(wr, 0)
log("weights equal across ranks:", bool(torch.equal(w, wr)))
dist.barrier(); log("barrier ok")
dist.destroy_process_group()
