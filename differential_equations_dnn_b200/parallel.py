"""Data parallelism over collocation rows and one-trial-per-GPU sweeps.

The reference is single-GPU.  Rows are independent (no batch-norm on the
hot path), so the only exchange per step is a SUM all-reduce of the flat FP32 buffer
[grad_theta (P) | loss (1)] -- 0.8 MB for dgm_net.DGM(2,1,128,3) -- over NCCL/NVLink;
every rank then applies the identical fused Adam update, so no broadcast is needed
(SURVEY 8e).  Each rank's kernels already scale by 1/B_global.

Hyper-parameter sweeps (optimize_heat_ray.py:133-203, batchsize_effect_heat.py:186-202)
are "replicas only": one process per GPU, no communication until the final gather.
"""
from __future__ import annotations

import os
import random

import torch
import torch.distributed as dist

_STATE = {"group": None, "enabled": False, "bglobal": None}


def enable_data_parallel(group=None, global_batch=None):
    """Shard every subsequent fused step over `group` (default: WORLD).

    `global_batch`: the (fixed) sum of the ranks' local row counts, if the caller knows it -- it is validated once,
    collectively, by the first step and saves the per-step count all-reduce.  Left None, every step all-reduces
    its local row count, so shards may be unequal and may change from step to step."""
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    _STATE.update(group=group, enabled=True, bglobal=None if global_batch is None else [int(global_batch), False])


def disable_data_parallel():
    _STATE.update(group=None, enabled=False, bglobal=None)


def is_enabled():
    return _STATE["enabled"]


def world_size():
    return dist.get_world_size(_STATE["group"]) if _STATE["enabled"] else 1


def rank():
    return dist.get_rank(_STATE["group"]) if _STATE["enabled"] else 0


def global_batch(B_local, device):
    """Sum of the ranks' local row counts.  Never cached per rank: with unequal or changing shards a per-rank cache
    would let one rank skip the collective another rank issues.  Either every rank all-reduces its count every
    step, or the caller declared a static global batch (`enable_data_parallel(global_batch=...)`), which every
    rank checks together on the first step."""
    if not _STATE["enabled"]:
        return B_local
    declared = _STATE["bglobal"]
    if declared is not None and declared[1]:
        return declared[0]
    t = torch.tensor([B_local], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_STATE["group"])
    Bg = int(t.item())
    if declared is not None:
        if Bg != declared[0]:
            raise RuntimeError(f"enable_data_parallel(global_batch={declared[0]}) but the ranks hold {Bg} rows")
        declared[1] = True
    return Bg


def reduce_step(launch, B_local, device):
    """Run `launch(B_global) -> [P+1] tensor` on this rank's rows and SUM it over ranks.  `device`: where the row
    count is all-reduced (the device of the step's tensors)."""
    if not _STATE["enabled"]:
        return launch(None)
    Bg = global_batch(B_local, device)
    out = launch(Bg)
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=_STATE["group"])
    return out


def sync_parameters(net):
    """Broadcast rank 0's flat parameter buffer: after this every replica holds identical weights whatever seeds
    the ranks constructed their networks with (the drivers call it once before the first step; the identical
    fused Adam update on the all-reduced gradient then keeps the replicas in step)."""
    if _STATE["enabled"]:
        dist.broadcast(net.flat_theta(), src=dist.get_global_rank(_STATE["group"], 0) if _STATE["group"] is not None else 0,
                       group=_STATE["group"])
    return net


def parameters_in_sync(net):
    """True when every rank holds the same weights (debug check: compares an all-reduced checksum)."""
    if not _STATE["enabled"]:
        return True
    th = net.flat_theta().double()
    mine = torch.stack([th.sum(), (th * th).sum()])
    lo, hi = mine.clone(), mine.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=_STATE["group"])
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=_STATE["group"])
    return bool(torch.equal(lo, hi))


def sampler_generator(device, base_seed=None):
    """The RNG the drivers draw collocation rows from.  Single process: None (torch's default generator, i.e. the
    reference's RNG stream).  Data parallel: a per-rank `torch.Generator(device)` seeded with base_seed + rank
    (base_seed defaults to torch.initial_seed()), so that the ranks draw DIFFERENT rows -- with one shared seed
    every rank would sample the same batch and N GPUs would do N times the work for no extra samples."""
    if not _STATE["enabled"]:
        return None
    gen = torch.Generator(device=device)
    gen.manual_seed((torch.initial_seed() if base_seed is None else int(base_seed)) + rank())
    return gen


def shard(t, dim=0):
    """This rank's contiguous block of rows (rank r gets [r*B/R, (r+1)*B/R))."""
    if not _STATE["enabled"]:
        return t
    R, r = world_size(), rank()
    B = t.shape[dim]
    lo, hi = (B * r) // R, (B * (r + 1)) // R
    return t.narrow(dim, lo, hi - lo)


# ---- independent trials, one per GPU (no Ray) -------------------------------------
def sample_search_space(n, seed=0):
    """n configs from the space of optimize_heat_ray.py:173-176:
    batch_size ~ randint[1,512), n_iters ~ randint[1000,50000), lrate ~ loguniform(1e-4,1e-1)."""
    import math
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        out.append({"batch_size": rng.randrange(1, 512), "n_iters": rng.randrange(1000, 50000),
                    "lrate": math.exp(rng.uniform(math.log(1e-4), math.log(1e-1)))})
    return out


def run_trials(objective, configs, seed=None):
    """Round-robin `configs` over the ranks (one GPU each, zero communication while
    training), gather `{config, loss}` records on every rank.  `objective(config)`
    returns the final loss, like objectiveRay -> session.report (optimize_heat_ray.py:157).
    `seed`: trial i starts from torch.manual_seed(seed + i), so a trial's result does not depend on how many
    GPUs the sweep runs on or on which rank it lands (the kernels are deterministic)."""
    if dist.is_initialized():
        R, r = dist.get_world_size(), dist.get_rank()
    else:
        R, r = 1, 0
    mine = [(i, c) for i, c in enumerate(configs) if i % R == r]
    results = []
    for i, c in mine:
        if seed is not None:
            torch.manual_seed(seed + i)
        results.append({"trial": i, "config": c, "loss": float(objective(c)), "rank": r})
    if R > 1:
        gathered = [None] * R
        dist.all_gather_object(gathered, results)
        results = [x for part in gathered for x in part]
    results.sort(key=lambda d: d["trial"])
    return results


def best_trial(results):
    """tune.ResultGrid.get_best_result(metric='loss', mode='min') (optimize_heat_ray.py:199-201)."""
    return min(results, key=lambda d: d["loss"])


def init_from_env(backend=None):
    """torchrun-style init (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*); binds cuda:LOCAL_RANK."""
    if dist.is_initialized():
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend=backend)
