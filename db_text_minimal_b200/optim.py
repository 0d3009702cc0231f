"""FlatAdam -- the optimizer step of the reference (torch.optim.Adam(dbnet.parameters(), lr=0.005, amsgrad=False),
src/train.py:114-117,172) as ONE launch over the executor's flat buffers.

DBTextModel's backward already writes every gradient into one flat float32 buffer (models.py: _grad_views; the same buffer
the data-parallel all-reduce works on).  FlatAdam re-homes the parameters into a flat buffer with the same layout, keeps
exp_avg / exp_avg_sq flat, and calls dbb_adam_step on the four buffers.  The step counter is a device scalar, so a training
step that ends in FlatAdam.step() is CUDA-graph capturable without torch's `capturable=True` machinery.

Same update rule as torch.optim.Adam (bias-corrected, eps added after the square root, L2 weight decay folded into the
gradient); parameters that never receive a gradient (the unused backbone.fc / backbone.smooth) are left untouched, as in
torch."""
import torch

from . import _lib


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False):
        if amsgrad:
            raise ValueError("FlatAdam: amsgrad is not implemented (the reference trains with amsgrad=False)")
        if not hasattr(model, "_flat_offsets"):
            raise TypeError("FlatAdam works on a db_text_minimal_b200 DBTextModel")
        plist = model._param_list()
        used = [p for i, p in enumerate(plist) if not model._unused[i]]
        super().__init__(used, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.model = model
        dev = used[0].device
        _lib.require_cuda(used[0])
        n = model._flat_numel
        self.flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self._scratch_g = None
        with torch.no_grad():       # re-home every trained parameter into the flat buffer (state_dict is unaffected)
            for i, p in enumerate(plist):
                if model._unused[i]:
                    continue
                o = model._flat_offsets[i]
                view = self.flat_p[o:o + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view

    def _flat_grad(self):
        """The flat gradient buffer of the last backward (zero-copy), or a gathered copy when the gradients were replaced."""
        model = self.model
        plist = model._param_list()
        flat = getattr(model, "_last_flat_grad", None)
        first = next(i for i in range(len(plist)) if not model._unused[i])
        g0 = plist[first].grad
        if g0 is None:
            return None
        if flat is not None and g0.data_ptr() == flat.data_ptr() + 4 * model._flat_offsets[first] and \
                all(plist[i].grad is not None and plist[i].grad.data_ptr() == flat.data_ptr() + 4 * model._flat_offsets[i]
                    for i in range(len(plist)) if not model._unused[i]):
            return flat
        if self._scratch_g is None:
            self._scratch_g = torch.zeros_like(self.flat_p)
        for i, p in enumerate(plist):
            if model._unused[i]:
                continue
            o = model._flat_offsets[i]
            dst = self._scratch_g[o:o + p.numel()]
            if p.grad is None:
                dst.zero_()
            else:
                dst.copy_(p.grad.reshape(-1))
        return self._scratch_g

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        loss = closure() if closure is not None else None
        g = self._flat_grad()
        if g is None:
            return loss
        hp = self.param_groups[0]
        with torch.cuda.device(self.flat_p.device):
            _lib.check(_lib.lib().dbb_adam_step(self.flat_p.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr(),
                                                self.exp_avg_sq.data_ptr(), self.flat_p.numel(), float(hp["lr"]),
                                                float(hp["betas"][0]), float(hp["betas"][1]), float(hp["eps"]),
                                                float(hp["weight_decay"]), float(grad_scale), self.step_count.data_ptr(),
                                                _lib.stream_ptr()), "dbb_adam_step")
        return loss
