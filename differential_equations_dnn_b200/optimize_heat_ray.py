"""Drop-in for the reference's optimize_heat_ray.py without Ray / Optuna: the hyper-parameter search over
(batch_size, n_iters, lrate) of the heat solver as independent trials, ONE TRIAL PER GPU, no communication while
training (SURVEY 8e "independent trials", 8f N3).

`minimize_loss_dgm` is the reference's copy of the heat driver in which BOTH boundary tensors sit at x = 0
(optimize_heat_ray.py:101-102,114-115); `objectiveRay(config)` trains `MLP(2, 1, 128, 3)` (ReLU, as constructed
there, :143-146) with the config and returns the last training loss -- what the reference hands to
`session.report({"loss": ...})` (:157); `optimizeHeat` draws `num_samples` configs from the reference's search
space (:173-176) and returns the best one (:199-203).  The reference's Optuna sampler and AsyncHyperBand early
stopping are Ray components, not arithmetic of this repository: the draws here are plain random search with a
fixed seed (`parallel.sample_search_space`).

    torchrun --nproc-per-node 8 -m differential_equations_dnn_b200.optimize_heat_ray --num-samples 10
"""
import contextlib
import io
import json
import sys
import time

import torch

from . import heat, parallel
from .neural_networks import MLP

dgm_loss_func = heat.dgm_loss_func
exact_solution = heat.exact_solution
gridEvaluation = heat.gridEvaluation


def minimize_loss_dgm(net, iterations=1000, batch_size=32, lrate=1e-4, cuda_graph=True):
    """optimize_heat_ray.py:80-130: the heat driver with X_BD2 = [0, t]."""
    return heat.minimize_loss_dgm(net, iterations=iterations, batch_size=batch_size, lrate=lrate, cuda_graph=cuda_graph,
                                  xbd2_value=0.0)


def objectiveRay(config, cuda_graph=True, quiet=True):
    """optimize_heat_ray.py:133-157.  Returns the reported loss (the reference reports it to the Ray session)."""
    with (contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext()):
        net = MLP(input_dim=2, output_dim=1, hidden_size=128, num_layers=3).cuda()
        _, loss_dgm = minimize_loss_dgm(net, iterations=int(config["n_iters"]), batch_size=int(config["batch_size"]),
                                        lrate=float(config["lrate"]), cuda_graph=cuda_graph)
    return loss_dgm[-1]


def optimizeHeat(num_samples=10, seed=0, objective=None, trial_seed=1234):
    """optimize_heat_ray.py:160-203: `num_samples` trials, round-robin one per GPU (rank), best config by final loss.
    Returns (best_config, all trial records)."""
    configs = parallel.sample_search_space(num_samples, seed)
    results = parallel.run_trials(objective or objectiveRay, configs, seed=trial_seed)
    return parallel.best_trial(results)["config"], results


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-samples", type=int, default=10)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default=None)
    a = ap.parse_args(argv)
    parallel.init_from_env()
    t0 = time.perf_counter()
    best, results = optimizeHeat(a.num_samples, a.seed)
    wall = time.perf_counter() - t0
    if (not torch.distributed.is_initialized()) or torch.distributed.get_rank() == 0:
        rec = {"study": "optimize_heat_ray", "trials": results, "best": best, "wall_s": wall,
               "world_size": torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1}
        print(json.dumps(rec))
        if a.out:
            json.dump(rec, open(a.out, "w"), indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
