"""Drop-in for the hot path of the reference's simple_ode.py: y' = -y, y(0) = 2 on [0,1].

Same names / signatures / return types as simple_ode.py:35-132; the loss runs in
dgmk_ode_step (include/dgmk.h).
"""
import numpy as np
import torch

from . import autograd as ag
from . import parallel
from ._flat import DeferredOutput, deferred_forward
from ._loop import graphed_loop, print_progress
from .auxiliary_funs import fn_timer
from .heat import _device
from .optim import FusedAdam
from .sampler import PhiloxSampler


def exact_solution(t):
    """2 exp(-t) (simple_ode.py:35-38)."""
    return 2.0 * np.exp(-t)


def _calls(y, y0):
    """(net, t, t0) when y and y0 are recorded or traceable calls of ONE of our networks."""
    if isinstance(y, DeferredOutput) and isinstance(y0, DeferredOutput) and y.net is y0.net:
        return y.net, y.x, y0.x
    if isinstance(y, torch.Tensor) and isinstance(y0, torch.Tensor):
        a, b = ag.producer_of(y), ag.producer_of(y0)
        if a is not None and b is not None and a[0] is b[0]:
            return a[0], a[1], b[1]
    return None


def _require_calls(y, y0):
    call = _calls(y, y0)
    if call is None:
        raise ag.DgmkError("dgm_loss_func(y, y0, t, y_ic): y and y0 must be what net(t), net(t0) of ONE network of "
                           "this package returned (directly, or recorded under deferred_forward); there is no "
                           "torch-autograd or CPU fallback path")
    return call


def dgm_loss_func(y, y0, t, y_ic):
    """mean[(y' + y)^2 + (y(0) - y_ic)^2] (simple_ode.py:41-63).

    `y`, `y0` are what `net(t)`, `net(t0)` returned.  If they came from one of this
    package's networks (eagerly, or recorded under `deferred_forward`) the fused step
    kernel computes loss and parameter gradient in one go.  Anything else (tensors of a
    foreign network, or outputs post-processed before the call) raises DgmkError: the
    package has no torch-autograd fallback."""
    net, tt, tt0 = _require_calls(y, y0)
    return ag.OdeStepFn.apply(net, tt, tt0, y_ic, *ag.params_of(net))


@fn_timer
def minimize_loss_dgm(net, y_ic=2.0, iterations=1000, batch_size=32, lrate=1e-4, cuda_graph=False, sampler="torch"):
    """simple_ode.py:66-112: t ~ 1.01 U[0,1), Adam(lr); returns (net, list[float]).
    `cuda_graph=True` (single GPU): one captured iteration replayed (`_loop.graphed_loop`).
    sampler="philox": t comes from this library's on-device Philox sampler (`sampler.PhiloxSampler`; statistically
    equivalent draws, one launch) instead of torch.rand."""
    if sampler not in ("torch", "philox"):
        raise ValueError("sampler must be 'torch' or 'philox'")
    device = _device()
    parallel.sync_parameters(net)            # data parallel: rank 0's weights everywhere
    gen = parallel.sampler_generator(device)  # ... and per-rank rows (None on one GPU: the default RNG stream)
    graphed = cuda_graph and not parallel.is_enabled()
    optimizer = FusedAdam(net.parameters(), lr=lrate, capturable=graphed)
    y_ic = torch.ones([batch_size, 1], device=device) * y_ic
    t0 = torch.zeros([batch_size, 1], device=device)
    ps = PhiloxSampler(device) if sampler == "philox" else None
    tp = torch.empty([batch_size, 1], device=device) if ps is not None else None

    def draw(i=0):
        if ps is not None:
            return ps.uniform(tp, 0.0, 1.01, step_add=i)
        return 1.01 * torch.rand([batch_size, 1], device=device, generator=gen)

    if graphed:
        def step():
            t = draw()
            optimizer.zero_grad()
            with deferred_forward(net):
                y, y0 = net(t), net(t0)
            loss = dgm_loss_func(y, y0, t, y_ic)
            loss.backward()
            optimizer.step()
            return loss
        train_loss = graphed_loop(step, iterations, device, counter=None if ps is None else ps.step)
        print_progress(train_loss, lrate, parallel.rank())
        return net, train_loss
    losses = []
    for i in range(iterations):
        t = draw(i)
        optimizer.zero_grad()
        with deferred_forward(net):
            y, y0 = net(t), net(t0)
        loss = dgm_loss_func(y, y0, t, y_ic)
        loss.backward()
        optimizer.step()
        losses.append(loss.detach())
        if i % 100 == 0 and parallel.rank() == 0:
            print(f"Iteration: {i}, Loss: {loss.item()}, LR: {optimizer.param_groups[0]['lr']}")
    return net, (torch.stack(losses).cpu().tolist() if losses else [])


def gridEvaluation(net, nodes=10):
    """net on `nodes` points of [0,1] (simple_ode.py:115-132), one batched launch."""
    t = torch.linspace(0, 1.0, nodes, dtype=torch.float64).float().reshape(-1, 1).to(_device())
    with torch.no_grad():
        return net(t)[:, 0].double().cpu().numpy()
