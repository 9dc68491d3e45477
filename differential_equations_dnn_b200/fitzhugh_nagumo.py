"""Drop-in for the hot path of the reference's fitzhugh_nagumo.py (two-variable ODE system).

Same names / signatures as fitzhugh_nagumo.py:38-177; the loss runs in dgmk_fhn_step.
"""
import numpy as np
import torch

from . import autograd as ag
from . import parallel
from ._flat import deferred_forward
from ._loop import graphed_loop, print_progress
from .auxiliary_funs import fn_timer
from .heat import _device
from .optim import FusedAdam
from .sampler import PhiloxSampler
from .simple_ode import _require_calls

IEXT, ALPHA, BETA, TAU = 0.5, 0.7, 0.8, 2.5  # fitzhugh_nagumo.py:69-70


def FZNFun(y, t, I, alpha, beta, tau):
    """Right-hand side for scipy.odeint (fitzhugh_nagumo.py:38-50)."""
    return np.array([y[0] - y[0] ** 3 / 3 - y[1] + I, (y[0] + alpha - beta * y[1]) / tau])


def dgm_loss_func(y, y0, t, y_ic):
    """mean(r_Y^2) + mean(r_W^2) + mean((y0 - y_ic)^2), r_Y = Y' + Y^3/3 + W - I - Y,
    r_W = W' + (beta W - alpha - Y)/tau (fitzhugh_nagumo.py:53-97; the last mean runs
    over 2B elements)."""
    net, tt, tt0 = _require_calls(y, y0)
    return ag.FhnStepFn.apply(net, tt, tt0, y_ic, *ag.params_of(net))


@fn_timer
def minimize_loss_dgm(net, y_ic, iterations=1000, batch_size=32, lrate=1e-4, sampler="grid", cuda_graph=False):
    """fitzhugh_nagumo.py:100-156.  sampler="grid" is the shipped one: `batch_size` (<=200)
    distinct nodes of a 200-point grid on [0,30] (:123-133); sampler="uniform" is the
    commented-out `30.01 * rand` (:129), the only one that scales past 200 rows; sampler="philox" is that uniform
    draw from this library's on-device Philox sampler (`sampler.PhiloxSampler`, one launch).
    `cuda_graph=True` (single GPU): one captured iteration replayed (`_loop.graphed_loop`)."""
    if sampler not in ("grid", "uniform", "philox"):
        raise ValueError("sampler must be 'grid', 'uniform' or 'philox'")
    device = _device()
    parallel.sync_parameters(net)            # data parallel: rank 0's weights everywhere
    gen = parallel.sampler_generator(device)  # ... and per-rank rows (None on one GPU: the default RNG stream)
    graphed = cuda_graph and not parallel.is_enabled()
    optimizer = FusedAdam(net.parameters(), lr=lrate, capturable=graphed)
    t0 = torch.zeros([batch_size, 1], device=device)
    num_samples = 200
    T = torch.linspace(0.0, 30.0, steps=num_samples, device=device)
    prob = torch.full((num_samples,), 1.0 / num_samples, device=device)

    ps = PhiloxSampler(device) if sampler == "philox" else None
    tp = torch.empty([batch_size, 1], device=device) if ps is not None else None

    def sample(i=0):
        if ps is not None:
            return ps.uniform(tp, 0.0, 30.01, step_add=i)
        if sampler == "grid":
            return T[prob.multinomial(num_samples=batch_size, replacement=False, generator=gen)].reshape(-1, 1)
        return 30.01 * torch.rand([batch_size, 1], device=device, generator=gen)

    if graphed:
        def step():
            t = sample()
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
        t = sample(i)
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
    """net on `nodes` points of [0,30] -> [nodes, 2] (fitzhugh_nagumo.py:159-177)."""
    t = torch.linspace(0.0, 30.0, nodes, dtype=torch.float64).float().reshape(-1, 1).to(_device())
    net.eval()
    with torch.no_grad():
        return net(t).double().cpu().numpy()
