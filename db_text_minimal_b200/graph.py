"""Whole-step CUDA-graph capture of the DB training step (forward + DBLoss + backward + optimizer).

The executor enqueues ~280 kernels per step; replaying them as one graph removes the per-launch host cost, which matters
when the caller synchronises every step (the reference reads the loss each iteration, src/train.py:188-201).
Everything the step touches is capturable by construction: the library never synchronises or allocates, tensor maps are
kernel arguments, and PyTorch's graph memory pool keeps workspace addresses stable across replays."""
import torch


class GraphedTrainStep:
    """step(img, gts) -> loss tensor (static, overwritten by every replay).

    Data-parallel use: the gradient all-reduces issued by dist.GradSync during backward are NCCL kernels on NCCL's own
    stream, forked from / joined to the capturing stream by events, so they become nodes of the same graph; pass
    capture_error_mode="thread_local" so that the process group's watchdog thread cannot invalidate the capture."""

    def __init__(self, model, criterion, optimizer, img_shape, gts_shape, device, warmup=3, capture_error_mode="global"):
        self.capture_error_mode = capture_error_mode
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.img = torch.zeros(img_shape, dtype=torch.float32, device=device)
        self.gts = torch.zeros(gts_shape, dtype=torch.float32, device=device)
        self.loss = None
        self.graph = None
        self._warmup = warmup

    def _eager(self):
        self.optimizer.zero_grad(set_to_none=True)
        losses = self.criterion(self.model(self.img), self.gts)
        losses[-1].backward()
        self.optimizer.step()
        return losses[-1].detach()

    def capture(self, img, gts):
        self.img.copy_(img); self.gts.copy_(gts)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self._warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib
        L = _lib.lib()
        self.graph = torch.cuda.CUDAGraph()
        n0 = L.dbb_launch_count()
        with torch.cuda.graph(self.graph, capture_error_mode=self.capture_error_mode):
            self.loss = self._eager()
        self.launches_per_replay = int(L.dbb_launch_count() - n0)     # kernel nodes of this library inside the graph
        return self

    def replay(self):
        """Run one step on whatever is currently in self.img / self.gts."""
        self.graph.replay()
        return self.loss

    def __call__(self, img, gts):
        self.img.copy_(img, non_blocking=True)
        self.gts.copy_(gts, non_blocking=True)
        return self.replay()
