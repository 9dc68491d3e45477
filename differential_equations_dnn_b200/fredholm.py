"""Drop-in for the hot path of the reference's fredholm.py:
y(x) = sin x + int_0^{pi/2} sin x cos t y(t) dt, Monte-Carlo integral with k fresh nodes
per point per step (fredholm.py:47-74).  Same names / signatures; the loss runs in
dgmk_fredholm_step.
"""
import numpy as np
import torch

from . import autograd as ag
from . import parallel
from ._flat import FlatParamModule
from ._loop import graphed_loop, print_progress
from .auxiliary_funs import fn_timer
from .heat import _device
from .optim import FusedAdam
from .sampler import PhiloxSampler


def exact_solution(x):
    """2 sin x (fredholm.py:40-44)."""
    return 2.0 * np.sin(x)


def draw_nodes(x, k, generator=None):
    """The k Monte-Carlo node sets, in the order the reference's loop draws them
    (`pi/2 * rand_like(x)`, fredholm.py:66-67) -> [k, B, 1].  `generator`: the per-rank sampler under data
    parallelism (parallel.sampler_generator); None = torch's default stream, like the reference."""
    if generator is None:
        return torch.stack([np.pi / 2.0 * torch.rand_like(x) for _ in range(k)])
    return torch.stack([np.pi / 2.0 * torch.rand(x.shape, device=x.device, generator=generator) for _ in range(k)])


def dgm_loss_func(net, x, k=50, nodes=None):
    """mean[(net(x) - sin x - (pi/2k) sum_j sin x cos t_j net(t_j))^2] (fredholm.py:47-74).
    `nodes` ([k,B,1]) lets a caller supply the draws (parity tests, data-parallel
    shards); by default they are drawn here exactly like the reference does."""
    if nodes is None:
        nodes = draw_nodes(x.detach(), k)
    if not isinstance(net, FlatParamModule):
        raise ag.DgmkError("dgm_loss_func needs one of this package's networks: there is no torch-autograd or CPU "
                           "fallback path")
    return ag.FredholmStepFn.apply(net, x, nodes, *ag.params_of(net))


@fn_timer
def minimize_loss_dgm(net, y_ic=2.0, iterations=1000, batch_size=32, lrate=1e-4, k=50, cuda_graph=False, sampler="torch"):
    """fredholm.py:77-117 (y_ic is accepted and unused, as there).
    `cuda_graph=True` (single GPU): one captured iteration replayed (`_loop.graphed_loop`).
    sampler="philox": the points and the k node sets come from this library's on-device Philox sampler
    (`sampler.PhiloxSampler`): 2 launches per step instead of the reference's 2 k + 3 (k rand_like, k multiplies, a stack);
    statistically equivalent draws, not torch's stream."""
    if sampler not in ("torch", "philox"):
        raise ValueError("sampler must be 'torch' or 'philox'")
    device = _device()
    parallel.sync_parameters(net)            # data parallel: rank 0's weights everywhere
    gen = parallel.sampler_generator(device)  # ... and per-rank rows (None on one GPU: the default RNG stream)
    graphed = cuda_graph and not parallel.is_enabled()
    optimizer = FusedAdam(net.parameters(), lr=lrate, capturable=graphed)
    ps = PhiloxSampler(device) if sampler == "philox" else None
    if ps is not None:
        tp = torch.empty([batch_size, 1], device=device)
        nodes_p = torch.empty([k, batch_size, 1], device=device)

    def draw(i=0):   # -> (points, nodes or None)
        if ps is not None:
            ps.uniform(tp, 0.0, np.pi / 2.0, stream_id=0, step_add=i)
            ps.uniform(nodes_p, 0.0, np.pi / 2.0, stream_id=1, step_add=i)
            return tp, nodes_p
        t = np.pi / 2.0 * torch.rand([batch_size, 1], device=device, generator=gen)
        return t, (None if gen is None else draw_nodes(t, k, gen))

    if graphed:
        def step():
            t, nodes = draw()
            optimizer.zero_grad()
            loss = dgm_loss_func(net, t, k, nodes=nodes)
            loss.backward()
            optimizer.step()
            return loss
        train_loss = graphed_loop(step, iterations, device, counter=None if ps is None else ps.step)
        print_progress(train_loss, lrate, parallel.rank())
        return net, train_loss
    losses = []
    for i in range(iterations):
        t, nodes = draw(i)
        optimizer.zero_grad()
        loss = dgm_loss_func(net, t, k, nodes=nodes)
        loss.backward()
        optimizer.step()
        losses.append(loss.detach())
        if i % 100 == 0 and parallel.rank() == 0:
            print(f"Iteration: {i}, Loss: {loss.item()}, LR: {optimizer.param_groups[0]['lr']}")
    return net, (torch.stack(losses).cpu().tolist() if losses else [])


def gridEvaluation(net, nodes=10, y_ic=2.0):
    """net on `nodes` points of [0, pi/2] (fredholm.py:120-138)."""
    t = torch.linspace(0, np.pi / 2.0, nodes, dtype=torch.float64).float().reshape(-1, 1).to(_device())
    with torch.no_grad():
        return net(t)[:, 0].double().cpu().numpy()
