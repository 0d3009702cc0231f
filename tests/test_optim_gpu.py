"""f-2: FlatAdam (one launch over the flat parameter / gradient buffers) against torch.optim.Adam, the optimizer of
src/train.py:114-117 (lr 0.005, betas (0.9, 0.999), eps 1e-8, no weight decay, amsgrad off)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _used(model):
    return [p for i, p in enumerate(model._param_list()) if not model._unused[i]]


@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_flat_adam_matches_torch_adam_on_given_gradients(wd):
    from db_text_minimal_b200.models import DBTextModel
    from db_text_minimal_b200.optim import FlatAdam
    torch.manual_seed(3)
    model = DBTextModel(pretrained=False).cuda().train()
    ref_params = [p.detach().clone().requires_grad_(True) for p in _used(model)]
    ref = torch.optim.Adam(ref_params, lr=0.005, weight_decay=wd)
    opt = FlatAdam(model, lr=0.005, weight_decay=wd)
    sd_before = {k: v.clone() for k, v in model.state_dict().items()}
    g = torch.Generator(device="cuda").manual_seed(5)
    for step in range(4):
        for p, q in zip(_used(model), ref_params):
            gr = torch.randn(p.shape, device="cuda", generator=g) * (10.0 ** (step - 2))
            p.grad = gr
            q.grad = gr.clone()
        opt.step()
        ref.step()
    worst = 0.0
    for p, q in zip(_used(model), ref_params):
        worst = max(worst, float((p.detach() - q.detach()).abs().max()))
        # tolerance: a few float32 ulps of a 0.005-sized step on O(1) parameters (fused-multiply ordering differs)
        assert torch.allclose(p.detach(), q.detach(), rtol=2e-6, atol=2e-7), worst
    assert int(opt.step_count.item()) == 4
    # parameters that never get a gradient are untouched; state_dict keys / shapes are unchanged by the re-homing
    sd_after = model.state_dict()
    assert list(sd_after.keys()) == list(sd_before.keys())
    for k in sd_after:
        if k.startswith("backbone.fc.") or k.startswith("backbone.smooth."):
            assert torch.equal(sd_after[k], sd_before[k])


def test_flat_adam_steps_on_the_backward_buffer_without_a_copy():
    from db_text_minimal_b200.models import DBTextModel
    from db_text_minimal_b200.losses import DBLoss
    from db_text_minimal_b200.optim import FlatAdam
    from db_text_minimal_b200 import synth
    torch.manual_seed(0)
    model = DBTextModel(pretrained=False).cuda().train()
    opt = FlatAdam(model, lr=0.005)
    crit = DBLoss(alpha=1.0, beta=10.0, reduction="mean", negative_ratio=3)
    img = synth.images(2, 64, 64, seed=1).cuda()
    gts = torch.from_numpy(synth.gt_maps(2, 64, 64, seed=1)).cuda()
    opt.zero_grad(set_to_none=True)
    crit(model(img), gts)[-1].backward()
    assert opt._flat_grad() is model._last_flat_grad                  # zero-copy: the executor's own buffer
    ref_params = [p.detach().clone().requires_grad_(True) for p in _used(model)]
    for q, p in zip(ref_params, _used(model)):
        q.grad = p.grad.detach().clone()
    ref = torch.optim.Adam(ref_params, lr=0.005)
    opt.step()
    ref.step()
    for p, q in zip(_used(model), ref_params):
        assert torch.allclose(p.detach(), q.detach(), rtol=2e-6, atol=2e-7)
    # the model runs on the re-homed parameters
    out = model(img)
    assert torch.isfinite(out).all()


def test_flat_adam_state_dict_resumes_moments_and_step():
    """optimizer.state_dict() / load_state_dict() must carry the flat moments and the step count, and a model
    that was moved after the optimizer was built must be detected."""
    from db_text_minimal_b200.models import DBTextModel
    from db_text_minimal_b200.optim import FlatAdam
    torch.manual_seed(3)
    g = torch.Generator(device="cuda").manual_seed(7)

    def grads(model):
        for p in _used(model):
            p.grad = torch.randn(p.shape, device="cuda", generator=g) * 1e-2

    a = DBTextModel(pretrained=False).cuda().train()
    oa = FlatAdam(a, lr=0.005)
    for _ in range(3):
        grads(a); oa.step()
    sd_m, sd_o = {k: v.clone() for k, v in a.state_dict().items()}, oa.state_dict()
    assert "flat" in sd_o and int(sd_o["flat"]["step"]) == 3
    b = DBTextModel(pretrained=False).cuda().train()
    b.load_state_dict(sd_m)
    ob = FlatAdam(b, lr=0.005)
    ob.load_state_dict(sd_o)
    gen_state = g.get_state()
    grads(a); oa.step()
    g.set_state(gen_state)
    grads(b); ob.step()
    for p, q in zip(_used(a), _used(b)):
        assert torch.equal(p.detach(), q.detach())
    assert int(ob.step_count.item()) == 4
    # re-homing is checked: a parameter swapped out from under the optimizer raises instead of training a stale copy
    first = _used(b)[0]
    first.data = first.data.clone()
    ob._step_calls = 0
    with pytest.raises(RuntimeError):
        grads(b); ob.step()
