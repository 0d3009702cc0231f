"""Data-parallel semantics on real GPUs over NCCL: every rank computes the loss terms and gradients of its
own shard; after GradSync's all-reduce the gradients equal the MEAN of the per-shard single-GPU gradients.  Needs two
GPUs (skipped otherwise; the gloo CPU test in tests/test_abi_cpu.py covers the host logic everywhere)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["DBB_ROOT"])
from db_text_minimal_b200 import DBLoss, DBTextModel, synth
from db_text_minimal_b200.dist import GradSync
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
torch.manual_seed(100 + rank)                      # DIFFERENT seeds on purpose: GradSync must broadcast rank 0's replica
model = DBTextModel(pretrained=False).cuda().train()
sync = GradSync(model)
ref = [p.detach().clone() for p in model.parameters()]
gathered = [torch.empty_like(ref[0]) for _ in range(world)]
dist.all_gather(gathered, ref[0])
assert all(torch.equal(g, gathered[0]) for g in gathered), "parameters were not broadcast"
crit = DBLoss(alpha=1.0, beta=10.0, reduction="none", negative_ratio=3)
n, s = 2, 128
img = synth.images(n, s, s, seed=10 + rank).cuda()
gts = torch.from_numpy(synth.gt_maps(n, s, s, seed=10 + rank)).cuda()
# 1) this rank's shard without synchronisation (single-GPU gradients of the shard)
model._segment_hook = None
crit(model(img), gts)[-1].backward()
own = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
# 2) the same step with the all-reduce
model.zero_grad(set_to_none=True)
model._segment_hook = sync
crit(model(img), gts)[-1].backward()
torch.cuda.synchronize()
worst = 0.0
for k, p in model.named_parameters():
    if p.grad is None:
        continue
    parts = [torch.empty_like(own[k]) for _ in range(world)]
    dist.all_gather(parts, own[k])
    mean = torch.stack(parts).mean(0)
    scale = float(mean.abs().max()) + 1e-30
    err = float((p.grad - mean).abs().max()) / scale
    worst = max(worst, err)
    assert err <= 1e-5, (k, err)
sync.sync_buffers(model)
rm = model.backbone.bn1.running_mean.detach().clone()
parts = [torch.empty_like(rm) for _ in range(world)]
dist.all_gather(parts, rm)
assert all(torch.equal(q, parts[0]) for q in parts)
if rank == 0:
    print("DP_PARITY_OK worst", worst)
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_allreduced_gradients_are_the_mean_of_the_shard_gradients(tmp_path):
    script = tmp_path / "dp_worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, DBB_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DP_PARITY_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
