"""world_size-2 gloo test of the data-parallel logic (parallel.reduce_step / shard):
each rank runs the LOCAL step on its shard scaled by 1/B_global and the SUM all-reduce of
[grad | loss] must equal the single-process result on the concatenated batch.  The local
step is the oracle here (CPU); on the GPU the same function wraps kernels.heat_step."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from differential_equations_dnn_b200 import parallel
    from oracle import jets_np as jn
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "heat_dgm_h32l1.npz")))
    parallel.enable_data_parallel()
    keys = ("X", "X0", "XBD1", "XBD2", "x_bd1", "x_bd2")
    # unequal shards on purpose: 64 rows -> 25 + 39
    cut = 25
    mine = [g[k][:cut] if rank == 0 else g[k][cut:] for k in keys]

    def launch(Bg):
        B = mine[0].shape[0]
        loss, grad = jn.heat_step(g["spec"], g["theta"].astype(np.float64), *[m.astype(np.float64) for m in mine])
        # oracle normalises by its local B; the kernels normalise by B_global
        return torch.from_numpy(np.concatenate([grad, [loss]]) * (B / Bg))

    out = parallel.reduce_step(launch, mine[0].shape[0], device=torch.device("cpu"))
    assert parallel.global_batch(mine[0].shape[0], torch.device("cpu")) == 64
    sh = parallel.shard(torch.arange(10))
    # replicas: different construction seeds per rank -> rank 0's weights after sync_parameters; per-rank sampler
    from differential_equations_dnn_b200 import dgm_net
    torch.manual_seed(100 + rank)
    net = dgm_net.DGM(2, 1, 8, 1)
    assert not parallel.parameters_in_sync(net)
    parallel.sync_parameters(net)
    assert parallel.parameters_in_sync(net)
    torch.manual_seed(5)   # the SAME seed on every rank (what a launcher typically does)
    gen = parallel.sampler_generator(torch.device("cpu"))
    draw = torch.rand(4, generator=gen)
    # a declared static global batch is validated collectively once, then used without a collective
    parallel.enable_data_parallel(global_batch=64)
    assert parallel.global_batch(mine[0].shape[0], torch.device("cpu")) == 64
    assert parallel.global_batch(mine[0].shape[0], torch.device("cpu")) == 64
    parallel.enable_data_parallel(global_batch=63)
    try:
        parallel.global_batch(mine[0].shape[0], torch.device("cpu"))
        bad = False
    except RuntimeError:
        bad = True
    assert bad
    q.put((rank, out.numpy(), sh.numpy(), net.flat_theta().numpy().copy(), draw.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_sum_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "heat_dgm_h32l1.npz")))
    for rank, out, sh, th, draw in res:
        assert abs(out[-1] - float(g["loss_f64"])) < 1e-12 * abs(float(g["loss_f64"]))
        assert np.linalg.norm(out[:-1] - g["grad_f64"]) / np.linalg.norm(g["grad_f64"]) < 1e-10
    assert np.array_equal(res[0][1], res[1][1])           # every rank holds the same reduced buffer
    assert np.array_equal(res[0][3], res[1][3])           # sync_parameters: rank 0's weights on every rank
    assert not np.array_equal(res[0][4], res[1][4])       # per-rank sampler: different rows from one shared seed
    assert list(res[0][2]) == [0, 1, 2, 3, 4] and list(res[1][2]) == [5, 6, 7, 8, 9]
