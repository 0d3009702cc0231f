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
        if len(self.param_groups) != 1:
            raise ValueError("FlatAdam keeps one flat buffer: exactly one parameter group is supported")
        self.model = model
        dev = used[0].device
        _lib.require_cuda(used[0])
        n = model._flat_numel
        self.flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self._scratch_g = None
        self._step_calls = 0
        with torch.no_grad():       # re-home every trained parameter into the flat buffer (state_dict is unaffected)
            for i, p in enumerate(plist):
                if model._unused[i]:
                    continue
                o = model._flat_offsets[i]
                view = self.flat_p[o:o + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view

    # ------------------------------------------------------------------ checkpoint / resume (src/train.py:285-291 saves the model;
    # a resumed run must also continue the Adam moments and the bias-correction step count)
    def state_dict(self):
        sd = super().state_dict()
        sd["flat"] = {"exp_avg": self.exp_avg.detach().clone(), "exp_avg_sq": self.exp_avg_sq.detach().clone(),
                      "step": self.step_count.detach().clone(), "numel": int(self.flat_p.numel())}
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        flat = state_dict.pop("flat", None)
        super().load_state_dict(state_dict)
        if flat is None:
            raise ValueError("FlatAdam.load_state_dict: not a FlatAdam state (no flat moments)")
        if int(flat["numel"]) != int(self.flat_p.numel()):
            raise ValueError("FlatAdam.load_state_dict: flat layout mismatch")
        self.exp_avg.copy_(flat["exp_avg"]); self.exp_avg_sq.copy_(flat["exp_avg_sq"]); self.step_count.copy_(flat["step"])

    def _check_homes(self):
        """Every trained parameter must still live inside flat_p (model.to() / .half() / a second FlatAdam re-home them)."""
        model = self.model
        base, end = self.flat_p.data_ptr(), self.flat_p.data_ptr() + 4 * self.flat_p.numel()
        for i, p in enumerate(model._param_list()):
            if model._unused[i]:
                continue
            if p.data_ptr() != base + 4 * model._flat_offsets[i] or not (base <= p.data_ptr() < end):
                raise RuntimeError("FlatAdam: parameter %s no longer lives in the optimizer's flat buffer (the model was moved / "
                                   "cast / re-wrapped after the optimizer was built); rebuild the optimizer" % model._pnames[i])

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
        if self._step_calls % 64 == 0:          # (a pointer walk over the 107 tensors: ~50 us, so not on every step)
            self._check_homes()
        self._step_calls += 1
        hp = self.param_groups[0]
        with torch.cuda.device(self.flat_p.device):
            _lib.check(_lib.lib().dbb_adam_step(self.flat_p.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr(),
                                                self.exp_avg_sq.data_ptr(), self.flat_p.numel(), float(hp["lr"]),
                                                float(hp["betas"][0]), float(hp["betas"][1]), float(hp["eps"]),
                                                float(hp["weight_decay"]), float(grad_scale), self.step_count.data_ptr(),
                                                _lib.stream_ptr()), "dbb_adam_step")
        return loss
