"""Data-parallel gradient synchronisation for DBTextModel.

One process per GPU; every rank holds a full replica and its own shard of the batch (BatchNorm statistics stay
per-rank, as in the single-device reference).  The only exchange is the gradient all-reduce: the native backward
(csrc/net.cu) writes all 12,269,378 gradient elements into ONE flat float32 buffer, ordered
[head + FPN | layer4, layer3 | layer2, layer1, stem]; after each backward segment the finished slice is all-reduced
asynchronously (NCCL over NVLink / NVSwitch picks ring / tree / NVLS) while the next segment computes.
"""
import torch
import torch.distributed as dist


class GradSync:
    """Installs itself as ``model._segment_hook``; averages gradients across the process group."""

    def __init__(self, model, group=None, broadcast=True):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.works = []
        self.pending = []
        model._segment_hook = self if self.world > 1 else None
        self.backend = dist.get_backend(group) if dist.is_initialized() else None
        if broadcast and self.world > 1:
            self.broadcast_state(model)

    def broadcast_state(self, model, src=0):
        """Every replica starts from rank `src`'s parameters AND buffers (as DistributedDataParallel does at construction):
        replicas built without a common seed would otherwise diverge silently."""
        with torch.no_grad():
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t.data, src=src, group=self.group)

    def sync_buffers(self, model):
        """Average the BatchNorm running statistics over the ranks (each rank normalises with its own shard's statistics, as
        the single-device reference would; call this before evaluating or saving a checkpoint so that it does not carry
        rank 0's statistics only)."""
        if self.world == 1:
            return
        with torch.no_grad():
            for name, b in model.named_buffers():
                if b.is_floating_point():
                    if self.backend == "nccl":
                        dist.all_reduce(b.data, op=dist.ReduceOp.AVG, group=self.group)
                    else:
                        dist.all_reduce(b.data, op=dist.ReduceOp.SUM, group=self.group)
                        b.data.div_(self.world)

    def __call__(self, seg, flat, bounds):
        if bounds is not None:
            a, b = bounds
            if b > a:
                self.reduce_slice(flat[a:b])
        else:
            self.finish()

    def reduce_slice(self, t):
        if self.world == 1:
            return
        if self.backend == "nccl":
            self.works.append(dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
        else:   # gloo (CPU tests): no AVG
            self.works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self.pending.append(t)

    def finish(self):
        for w in self.works:
            w.wait()
        for t in self.pending:
            t.div_(self.world)
        self.works, self.pending = [], []


def shard_range(total, rank, world):
    """Contiguous shard of ``total`` independent units (images) owned by ``rank``."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
