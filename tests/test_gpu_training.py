"""Training on the kernels (g1 v0.5): with gradients required the FORWARD still runs the sm_100a kernels (same
values as inference, launches counted), the backward is the gradient of the differentiable restatement."""
import numpy as np
import pytest
import torch

import cases
from helpers import golden, rel_max, ro_module, sa_module
from slotformer_b200 import engine

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _grads(params):
    return [None if p.grad is None else p.grad.detach().clone() for p in params]


def _assert_same_grads(g_kernel, g_eager, tol=1e-4):
    """Both backward passes differentiate the SAME restatement at the SAME inputs with the SAME upstream gradient
    (the loss is linear in the output), so they agree up to the summation order of the atomics in torch's backward
    kernels.  The error is measured against the largest gradient entry of the whole set: a gradient that is
    analytically zero (project_q.0.bias only shifts every slot's logit by the same amount) is pure rounding noise."""
    scale = max(float(b.abs().max()) for b in g_eager if b is not None)
    for a, b in zip(g_kernel, g_eager):
        assert (a is None) == (b is None)
        if a is not None:
            assert float((a - b).abs().max()) < tol * max(scale * 1e-3, float(b.abs().max())), (a.shape,)


def test_slot_attention_grad_mode_forward_is_the_kernel_and_grads_match_the_restatement():
    c, w, feats, slots = cases.sa_case('sa_cfg1')
    m = sa_module(c, w, DEV, mask=True).train()
    f = torch.from_numpy(feats).to(DEV).requires_grad_(True)
    s0 = torch.from_numpy(slots).to(DEV).requires_grad_(True)
    params = list(m.parameters())
    with torch.no_grad():
        ref_out, ref_mask = m(f, s0)
    n0 = engine.launch_count()
    out, mask = m(f, s0)
    assert engine.launch_count() > n0                      # the kernels ran
    assert out.requires_grad and not mask.requires_grad
    assert torch.equal(out, ref_out) and torch.equal(mask, ref_mask)
    gw = torch.randn(out.shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5))
    ((out * gw).sum() + 0. * mask.sum()).backward()
    g_kernel = _grads([f, s0] + params)
    for t in [f, s0] + params:
        t.grad = None
    e_out, _ = m._autograd_forward(f, s0, True)
    (e_out * gw).sum().backward()
    g_eager = _grads([f, s0] + params)
    assert rel_max(out.detach().cpu().numpy(), e_out.detach().cpu().numpy()) < 1e-3
    _assert_same_grads(g_kernel, g_eager)


def test_rollout_grad_mode_forward_is_the_kernel_in_eval_and_restatement_with_dropout():
    c, w, hist = cases.ro_case('ro_cfg2')
    m = ro_module(c, w, DEV, enc_t_pe=golden('ro_cfg2')['enc_t_pe']).eval()
    x = torch.from_numpy(hist).to(DEV).requires_grad_(True)
    with torch.no_grad():
        ref = m(x, 4)
    n0 = engine.launch_count()
    out = m(x, 4)                                          # eval(): no dropout -> kernel forward with autograd
    assert engine.launch_count() == n0 + 1 and out.requires_grad
    assert torch.equal(out, ref)
    gw = torch.randn(out.shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(6))
    (out * gw).sum().backward()
    params = [p for p in m.parameters() if p.requires_grad]
    g_kernel = _grads([x] + params)
    for t in [x] + params:
        t.grad = None
    e_out = m._autograd_forward(x, 4)
    (e_out * gw).sum().backward()
    _assert_same_grads(g_kernel, _grads([x] + params))
    m.train()                                              # dropout 0.1 active: the reference's training forward
    n0 = engine.launch_count()
    with pytest.warns(UserWarning, match='dropout'):
        from slotformer_b200 import autograd as ag
        ag._warned.discard('ro_dropout')
        d = m(x, 2)
    assert engine.launch_count() == n0 and d.requires_grad
