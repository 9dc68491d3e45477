"""Drop-in for the study in the reference's batchsize_effect_heat.py (:175-203): mean training-loss curve of the heat
solver per batch size 2^0 .. 2^10, `n_runs` runs of `n_iters` iterations each, `MLP(2, 1, 128, 3)`.

As shipped the reference's loop has two quirks, both reproduced by default (results must match the reference):
  * it passes `batch_size=64` to the driver whatever the loop variable says (:197), so all eleven curves are the
    batch-64 curve;
  * it builds ONE network before the loops (:181-184) and keeps training it across runs and batch sizes.
`fix_batch_size=True` trains with the loop's batch size; `fresh_net=True` re-initialises the network per run.
The eleven batch sizes are independent trials: under torchrun they are spread round-robin over the GPUs
(`parallel.run_trials`), one trial per GPU, no communication until the final gather.

    torchrun --nproc-per-node 8 -m differential_equations_dnn_b200.batchsize_effect_heat --n-iters 15000 --fix-batch-size --fresh-net
"""
import contextlib
import io
import json
import sys
import time

import numpy as np
import torch

from . import parallel
from .neural_networks import MLP
from .optimize_heat_ray import minimize_loss_dgm   # the same X_BD2 = [0, t] copy (batchsize_effect_heat.py:120,133)


def loss_curve(net, batch_size, n_iters, n_runs, fix_batch_size=False, fresh_net=False, cuda_graph=True):
    """Mean over `n_runs` runs of the loss trajectory (batchsize_effect_heat.py:191-202)."""
    running = np.zeros((n_runs, n_iters))
    for i in range(n_runs):
        if fresh_net:
            net = MLP(input_dim=2, output_dim=1, hidden_size=128, num_layers=3).cuda()
        _, loss = minimize_loss_dgm(net, iterations=n_iters, batch_size=batch_size if fix_batch_size else 64, lrate=1e-4,
                                    cuda_graph=cuda_graph)
        running[i] = loss
    return running.mean(axis=0)


def run_study(n_iters=15000, n_runs=5, n_batches=10, fix_batch_size=False, fresh_net=False, cuda_graph=True, trial_seed=1234):
    """-> [{"trial", "config": {"batch_size"}, "loss" (last point of the mean curve), "curve"}] for 2^0 .. 2^n_batches."""
    with contextlib.redirect_stdout(io.StringIO()):
        shared = None if fresh_net else MLP(input_dim=2, output_dim=1, hidden_size=128, num_layers=3).cuda()
    curves = {}

    def objective(cfg):
        with contextlib.redirect_stdout(io.StringIO()):
            c = loss_curve(shared, cfg["batch_size"], n_iters, n_runs, fix_batch_size, fresh_net, cuda_graph)
        curves[cfg["batch_size"]] = c
        return c[-1]
    configs = [{"batch_size": 2 ** i} for i in range(n_batches + 1)]
    results = parallel.run_trials(objective, configs, seed=trial_seed)
    for r in results:   # curves of the trials this rank ran (rank 0 holds all of them on one GPU)
        c = curves.get(r["config"]["batch_size"])
        if c is not None:
            r["curve_every_100"] = [float(v) for v in c[::100]]
    return results


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-iters", type=int, default=15000)
    ap.add_argument("--n-runs", type=int, default=5)
    ap.add_argument("--n-batches", type=int, default=10)
    ap.add_argument("--fix-batch-size", action="store_true")
    ap.add_argument("--fresh-net", action="store_true")
    ap.add_argument("--out", default=None)
    a = ap.parse_args(argv)
    parallel.init_from_env()
    torch.manual_seed(1234)
    t0 = time.perf_counter()
    results = run_study(a.n_iters, a.n_runs, a.n_batches, a.fix_batch_size, a.fresh_net)
    wall = time.perf_counter() - t0
    if (not torch.distributed.is_initialized()) or torch.distributed.get_rank() == 0:
        rec = {"study": "batchsize_effect_heat", "n_iters": a.n_iters, "n_runs": a.n_runs, "fix_batch_size": a.fix_batch_size,
               "fresh_net": a.fresh_net, "trials": results, "wall_s": wall,
               "world_size": torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1}
        print(json.dumps(rec))
        if a.out:
            json.dump(rec, open(a.out, "w"), indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
