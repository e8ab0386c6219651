"""Training on the kernels ("v0.5"): ``torch.autograd.Function`` wrappers whose FORWARD
is the sm_100a kernel.  The backward pass recomputes the operator with its differentiable PyTorch restatement
(the module's ``_autograd_forward``) from the saved inputs and takes that graph's gradient -- the reference
back-propagates through every Slot Attention iteration (savi.py:76-100, no stop-gradient) and through the whole
autoregressive chain (slotformer.py:113-124), and so does the restatement.  Forward values therefore come from
the same kernels in training and in inference; gradients are exact for the restatement, evaluated at inputs that
differ from the kernel's by its stated tolerance.  A hand-written backward is the next step.
"""
import warnings

import torch

_warned = set()


def warn_once(key, message):
    if key not in _warned:
        _warned.add(key)
        warnings.warn(message, stacklevel=3)


class KernelForward(torch.autograd.Function):
    """``apply(kernel_fn, eager_fn, n_inputs, n_diff_outputs, *inputs, *params)``.

    kernel_fn(*inputs) -> tensor or tuple (run under no_grad: the engine call);
    eager_fn(*inputs)  -> the same outputs through differentiable ops that read the live parameters;
    the first ``n_diff_outputs`` outputs are differentiable, the rest (e.g. the seg mask) are not."""

    @staticmethod
    def forward(ctx, kernel_fn, eager_fn, n_inputs, n_diff, *tensors):
        inputs = tensors[:n_inputs]
        with torch.no_grad():
            out = kernel_fn(*inputs)
        outs = out if isinstance(out, tuple) else (out,)
        ctx.eager_fn, ctx.n_inputs, ctx.n_diff, ctx.single = eager_fn, n_inputs, n_diff, not isinstance(out, tuple)
        ctx.save_for_backward(*tensors)
        ctx.mark_non_differentiable(*outs[n_diff:])
        return out

    @staticmethod
    def backward(ctx, *grad_outs):
        tensors = ctx.saved_tensors
        n_in = ctx.n_inputs
        inputs = [t.detach().requires_grad_(ctx.needs_input_grad[4 + i]) for i, t in enumerate(tensors[:n_in])]
        params = list(tensors[n_in:])
        with torch.enable_grad():
            out = ctx.eager_fn(*inputs)
        outs = (out if isinstance(out, tuple) else (out,))[:ctx.n_diff]
        wanted = [t for t in inputs if t.requires_grad] + [p for p in params if p.requires_grad]
        grads = torch.autograd.grad(outs, wanted, grad_outs[:ctx.n_diff], allow_unused=True) if wanted else ()
        it = iter(grads)
        g_in = [next(it) if t.requires_grad else None for t in inputs]
        g_par = [next(it) if p.requires_grad else None for p in params]
        return (None, None, None, None, *g_in, *g_par)
