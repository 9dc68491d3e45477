"""Drop-in for the hot path of the reference's heat.py: 1-D heat equation u_t = k u_xx.

`dgm_loss_func`, `minimize_loss_dgm`, `gridEvaluation`, `exact_solution` keep the
reference's names, argument order and return types (heat.py:36-172).  The loss is one
fused kernel sequence (include/dgmk.h: dgmk_heat_step) instead of four network
evaluations and two nested torch.autograd.grad sweeps.
"""
import numpy as np
import torch

from . import autograd as ag
from . import parallel
from ._flat import DeferredOutput, FlatParamModule, deferred_forward  # noqa: F401
from ._loop import graphed_loop, print_progress
from .auxiliary_funs import fn_timer
from .optim import FusedAdam
from .sampler import PhiloxSampler


def _device():
    if not torch.cuda.is_available():
        raise ag.DgmkError("no CUDA device: differential_equations_dnn_b200 has no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def exact_solution(k=1, nodes=10):
    """sin(x) exp(-k t) on the (t, x) grid [0,3] x [0,pi] (heat.py:36-47)."""
    t = np.linspace(0, 3, nodes)[:, None]
    x = np.linspace(0, np.pi, nodes)[None, :]
    return np.sin(x) * np.exp(-k * t)


def dgm_loss_func(net, x, x0, xbd1, xbd2, x_bd1, x_bd2):
    """mean[(u_t - k u_xx)^2 + (u(x,0) - sin x)^2 + (u(0,t) - x_bd1)^2 + (u(pi,t) - x_bd2)^2]
    with k = 1 (heat.py:50-95).  Returns a 0-dim tensor; `.backward()` fills `.grad`."""
    if not isinstance(net, FlatParamModule):
        raise ag.DgmkError("dgm_loss_func needs one of this package's networks (neural_networks.MLP / DGM, "
                           "dgm_net.DGM): there is no torch-autograd or CPU fallback path")
    return ag.HeatStepFn.apply(net, x, x0, xbd1, xbd2, x_bd1, x_bd2, 1.0, *ag.params_of(net))


def _minimize_graphed(net, iterations, batch_size, lrate, warmup=11, xbd2_value=torch.pi, sampler="torch"):
    """The training loop with one captured iteration replayed (`_loop.graphed_loop`): same RNG stream,
    same arithmetic as the eager loop (sampler="torch"), or the four operand blocks drawn by ONE launch of the
    on-device Philox sampler (sampler="philox", SURVEY 8f N2: 8 sampler launches per iteration become 1)."""
    device = _device()
    parallel.sync_parameters(net)            # data parallel: rank 0's weights everywhere
    gen = parallel.sampler_generator(device)  # ... and per-rank rows (None on one GPU: the default RNG stream)
    optimizer = FusedAdam(net.parameters(), lr=lrate, capturable=True)
    t0 = torch.zeros([batch_size, 1], device=device)
    xbd1 = torch.zeros([batch_size, 1], device=device)
    xbd2x = torch.ones([batch_size, 1], device=device) * xbd2_value
    xbd2y = torch.zeros([batch_size, 1], device=device)

    ps = PhiloxSampler(device) if sampler == "philox" else None
    if ps is not None:
        PX, PX0, PB1, PB2 = (torch.empty([batch_size, 2], device=device) for _ in range(4))

    def step():
        if ps is not None:
            ps.heat(PX, PX0, PB1, PB2, torch.pi, 3.0, xbd2_value)
            X, X0, X_BD1, X_BD2 = PX, PX0, PB1, PB2
        else:
            x = torch.pi * torch.rand([batch_size, 1], device=device, generator=gen)
            t = 3.0 * torch.rand([batch_size, 1], device=device, generator=gen)
            X = torch.cat([x, t], dim=1)
            X0 = torch.cat([x, t0], dim=1)
            X_BD1 = torch.cat([xbd1, t], dim=1)
            X_BD2 = torch.cat([xbd2x, t], dim=1)
        optimizer.zero_grad()
        loss = dgm_loss_func(net, X, X0, X_BD1, X_BD2, xbd1, xbd2y)
        loss.backward()
        optimizer.step()
        return loss

    train_loss = graphed_loop(step, iterations, device, warmup, counter=None if ps is None else ps.step)
    print_progress(train_loss, lrate, parallel.rank())
    return net, train_loss


@fn_timer
def minimize_loss_dgm(net, iterations=1000, batch_size=32, lrate=1e-4, cuda_graph=False, xbd2_value=torch.pi,
                      sampler="torch"):
    """The reference's training driver (heat.py:98-149): same sampler, same Adam
    defaults, returns (net, train_loss: list[float]).  Differences that do not change
    results: losses stay on the device and are read back once at the end (plus every
    100th for the progress print) instead of a host sync per step (heat.py:143); under
    `parallel.enable_data_parallel()` rank 0's weights are broadcast first, each rank draws its own
    rows from a per-rank generator (`parallel.sampler_generator`) and the gradient is all-reduced
    inside `dgm_loss_func`.  `cuda_graph=True` (single GPU) replays one captured
    iteration instead of launching it from Python (`_minimize_graphed`).  sampler="philox": the collocation points
    come from this library's on-device Philox sampler (one launch per step; `sampler.PhiloxSampler`) instead of
    torch.rand -- statistically equivalent draws, not torch's stream."""
    if sampler not in ("torch", "philox"):
        raise ValueError("sampler must be 'torch' or 'philox'")
    if cuda_graph and not parallel.is_enabled():
        return _minimize_graphed(net, iterations, batch_size, lrate, xbd2_value=xbd2_value, sampler=sampler)
    device = _device()
    parallel.sync_parameters(net)            # data parallel: rank 0's weights everywhere
    gen = parallel.sampler_generator(device)  # ... and per-rank rows (None on one GPU: the default RNG stream)
    optimizer = FusedAdam(net.parameters(), lr=lrate)
    t0 = torch.zeros([batch_size, 1], device=device)
    xbd1 = torch.zeros([batch_size, 1], device=device)
    xbd2x = torch.ones([batch_size, 1], device=device) * xbd2_value
    xbd2y = torch.zeros([batch_size, 1], device=device)
    losses = []
    ps = PhiloxSampler(device) if sampler == "philox" else None
    if ps is not None:
        X, X0, X_BD1, X_BD2 = (torch.empty([batch_size, 2], device=device) for _ in range(4))
    for i in range(iterations):
        if ps is not None:
            ps.heat(X, X0, X_BD1, X_BD2, torch.pi, 3.0, xbd2_value, step_add=i)
        else:
            x = torch.pi * torch.rand([batch_size, 1], device=device, generator=gen)
            t = 3.0 * torch.rand([batch_size, 1], device=device, generator=gen)
            X = torch.cat([x, t], dim=1)
            X0 = torch.cat([x, t0], dim=1)
            X_BD1 = torch.cat([xbd1, t], dim=1)
            X_BD2 = torch.cat([xbd2x, t], dim=1)
        optimizer.zero_grad()
        loss = dgm_loss_func(net, X, X0, X_BD1, X_BD2, xbd1, xbd2y)
        loss.backward()
        optimizer.step()
        losses.append(loss.detach())
        if i % 100 == 0 and parallel.rank() == 0:
            print(f"Iteration: {i}, Loss: {loss.item()}, LR: {optimizer.param_groups[0]['lr']}")
    train_loss = torch.stack(losses).cpu().tolist() if losses else []
    return net, train_loss


def gridEvaluation(net, nodes=10):
    """net on the nodes x nodes (t, x) grid (heat.py:152-172): one batched value-only
    launch instead of nodes^2 single-row forwards."""
    device = _device()
    t = torch.linspace(0, 3.0, nodes, dtype=torch.float64)
    x = torch.linspace(0, np.pi, nodes, dtype=torch.float64)
    T, Xg = torch.meshgrid(t, x, indexing="ij")
    pts = torch.stack([Xg.reshape(-1), T.reshape(-1)], 1).float().to(device)
    with torch.no_grad():
        y = net(pts)
    return y.reshape(nodes, nodes).double().cpu().numpy()
