"""torch.autograd glue between the drop-in modules / loss functions and the kernels.

Two seams:

S1  `net(x)` must stay differentiable twice in `x` and once in the parameters, because
    the UNMODIFIED reference losses call torch.autograd.grad(y, x, create_graph=True)
    on it (heat.py:73-85, simple_ode.py:54-58, fitzhugh_nagumo.py:74-84).  `JetFn`
    launches the jet kernels once (value, Jacobian, Hessian for a detached x) and
    `Link0`/`Link1` replay those through autograd's double-backward protocol with
    cheap einsums; `loss.backward()` then reaches `JetFn.backward` exactly once with the
    cotangents of all three outputs and launches the fused reverse pass.

S2  the package's own `dgm_loss_func`s go straight to the fused step kernels
    (`*StepFn`): one launch sequence yields the loss AND d loss / d theta; backward only
    scales and hands out views of the flat gradient.
"""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import kernels
from ._cabi import DgmkError


def _need_cuda(*ts):
    for t in ts:
        if isinstance(t, torch.Tensor) and not t.is_cuda:
            raise DgmkError("differential_equations_dnn_b200 runs on CUDA tensors only: there is no "
                            "CPU or PyTorch fallback path (move the module and inputs to a B200)")


def _f32c(t):
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _rows_like(t, B, cols, ref, name):
    """A per-row operand the reference would broadcast -- a Python number, a 0-dim / [1,cols] / [B,1] tensor
    (`(y0 - y_ic)`, simple_ode.py:62; `net(xbd1) - x_bd1`, heat.py:91-94) -- expanded to the contiguous
    [B, cols] float32 block the kernels index."""
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t, dtype=torch.float32, device=ref.device)
    _need_cuda(t)
    t = _f32c(t)
    if t.dim() > 2:
        raise DgmkError(f"{name}: expected a tensor broadcastable to [{B}, {cols}], got {list(t.shape)}")
    try:
        return t.expand(B, cols).contiguous()
    except RuntimeError:
        raise DgmkError(f"{name}: expected a tensor broadcastable to [{B}, {cols}], got {list(t.shape)}") from None


def _split_grads(net, flat_grad):
    """Views of a flat gradient, one per parameter; None for never-used parameters."""
    return tuple(flat_grad[off:off + n].view(p.shape) if live else None
                 for p, off, n, live in net.param_slices())


# =============================== S1: module forward ==================================
class JetFn(Function):
    """(params...) -> Y [B,o], J [B,o,d], Hs [B,o,d,d] at a detached x."""

    @staticmethod
    def forward(ctx, net, x, order, *params):
        theta = net.flat_theta()
        Y, J, Hs, stash = kernels.jet_forward(net.desc, theta, x, order)
        ctx.net, ctx.x, ctx.order, ctx.stash = net, x, order, stash
        ctx._dgmk_call = (net, x)
        B, o, d = x.shape[0], Y.shape[1], x.shape[1]
        if J is None:
            J = x.new_zeros(B, o, d)
        if Hs is None:
            Hs = x.new_zeros(B, o, d, d)
        return Y, J, Hs

    @staticmethod
    @once_differentiable
    def backward(ctx, gY, gJ, gHs):
        net, order = ctx.net, ctx.order
        gY = _f32c(gY) if gY is not None else None
        gJ = _f32c(gJ) if (gJ is not None and order >= 1) else None
        gHs = _f32c(gHs) if (gHs is not None and order >= 2) else None
        flat = kernels.jet_reverse(net.desc, net.flat_theta(), ctx.x, order, gY, gJ, gHs, ctx.stash)
        ctx.stash = None
        return (None, None, None) + _split_grads(net, flat)


class Link1(Function):
    """gx = sum_o gy[:,o] J[:,o,:]  -- what autograd.grad(y, x, gy) returns; its own
    backward supplies the second derivatives from Hs."""

    @staticmethod
    def forward(ctx, x, gy, J, Hs):
        ctx.save_for_backward(gy, J, Hs)
        return torch.einsum("bo,bod->bd", gy, J)

    @staticmethod
    def backward(ctx, ggx):
        gy, J, Hs = ctx.saved_tensors
        gx = torch.einsum("bd,bo,bode->be", ggx, gy, Hs)   # differentiable in Hs -> JetFn
        ggy = torch.einsum("bd,bod->bo", ggx, J)
        gJ = torch.einsum("bd,bo->bod", ggx, gy)
        return gx, ggy, gJ, None


class Link0(Function):
    """Identity on y that ties it to x: backward hands (x-gradient via Link1, gy)."""

    @staticmethod
    def forward(ctx, net, x, y, J, Hs):
        ctx.save_for_backward(x, J, Hs)
        ctx._dgmk_call = (net, x)
        return y.view_as(y)

    @staticmethod
    def backward(ctx, gy):
        x, J, Hs = ctx.saved_tensors
        return None, Link1.apply(x, gy, J, Hs), gy, None, None


def module_forward(net, x):
    """`net(x)` for x [B, d] (FlatParamModule.forward)."""
    _need_cuda(x, net.flat_theta())
    xd = _f32c(x)
    grad_on = torch.is_grad_enabled()
    params = [p for p, *_ in net.param_slices()]
    wants_param_grad = grad_on and any(p.requires_grad for p in params)
    if grad_on and x.requires_grad:
        order = int(net.jet_order)
        Y, J, Hs = JetFn.apply(net, xd, order, *params)
        return Link0.apply(net, x, Y, J, Hs)
    if wants_param_grad:
        Y, _, _ = JetFn.apply(net, xd, 0, *params)
        return Y
    return kernels.evaluate(net.desc, net.flat_theta(), xd)


# =============================== S2: fused steps =====================================
class _StepFn(Function):
    """Shared shape of the four fused steps: forward computes [grad | loss] in one
    kernel sequence; backward scales by the incoming cotangent."""

    @staticmethod
    def _run(B_local, launch, device):
        from . import parallel
        return parallel.reduce_step(launch, B_local, device)  # [P + 1] = grad_theta | loss

    @staticmethod
    def _finish(ctx, net, out):
        P = out.numel() - 1
        ctx.net = net
        ctx.flat_grad = out[:P]
        return out[P].clone()

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        net = ctx.net
        scaled = ctx.flat_grad * g
        nin = ctx.n_inputs
        return (None,) * nin + _split_grads(net, scaled)


class HeatStepFn(_StepFn):
    @staticmethod
    def forward(ctx, net, x, x0, xbd1, xbd2, x_bd1, x_bd2, kappa, *params):
        _need_cuda(x, x0, xbd1, xbd2)
        a = [_f32c(t) for t in (x, x0, xbd1, xbd2)]
        B = a[0].shape[0]
        a += [_rows_like(x_bd1, B, 1, a[0], "x_bd1"), _rows_like(x_bd2, B, 1, a[0], "x_bd2")]
        out = _StepFn._run(a[0].shape[0], lambda Bg: kernels.heat_step(net.desc, net.flat_theta(), *a, kappa=kappa,
                                                             B_global=Bg), a[0].device)
        ctx.n_inputs = 8
        return _StepFn._finish(ctx, net, out)


class OdeStepFn(_StepFn):
    @staticmethod
    def forward(ctx, net, t, t0, y_ic, *params):
        _need_cuda(t, t0)
        a = [_f32c(t), _f32c(t0)]
        a.append(_rows_like(y_ic, a[0].shape[0], net.desc.output_dim, a[0], "y_ic"))
        out = _StepFn._run(a[0].shape[0], lambda Bg: kernels.ode_step(net.desc, net.flat_theta(), *a, B_global=Bg), a[0].device)
        ctx.n_inputs = 4
        return _StepFn._finish(ctx, net, out)


class FhnStepFn(_StepFn):
    @staticmethod
    def forward(ctx, net, t, t0, y_ic, *params):
        _need_cuda(t, t0)
        a = [_f32c(t), _f32c(t0)]
        a.append(_rows_like(y_ic, a[0].shape[0], net.desc.output_dim, a[0], "y_ic"))
        out = _StepFn._run(a[0].shape[0], lambda Bg: kernels.fhn_step(net.desc, net.flat_theta(), *a, B_global=Bg), a[0].device)
        ctx.n_inputs = 4
        return _StepFn._finish(ctx, net, out)


class FredholmStepFn(_StepFn):
    @staticmethod
    def forward(ctx, net, x, nodes, *params):
        a = [_f32c(z) for z in (x, nodes)]
        out = _StepFn._run(a[0].shape[0], lambda Bg: kernels.fredholm_step(net.desc, net.flat_theta(), *a, B_global=Bg), a[0].device)
        ctx.n_inputs = 3
        return _StepFn._finish(ctx, net, out)


def params_of(net):
    return [p for p, *_ in net.param_slices()]


def producer_of(y):
    """Recover (net, x) from a tensor produced by `module_forward` with x.requires_grad
    (the Link0 node), so that `dgm_loss_func(y, y0, t, y_ic)` -- which receives network
    OUTPUTS, simple_ode.py:41 -- can still take the fused path."""
    fn = getattr(y, "grad_fn", None)
    return getattr(fn, "_dgmk_call", None) if fn is not None else None
