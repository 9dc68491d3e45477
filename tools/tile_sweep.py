"""Resident-tile step vs layer-wise path over batch sizes (crossover for the dispatch rule in CudaBackend::tile_step)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from differential_equations_dnn_b200 import dgm_net, neural_networks, kernels as K, _cabi
lib = _cabi.load()
lib.dgmk_set_tile_dispatch(1) if hasattr(lib, "dgmk_set_tile_dispatch") else None
def bench(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
cases = [("heat", 32, 1), ("heat", 64, 3), ("heat", 50, 3), ("fhn", 64, 2), ("ode", 32, 1), ("heat", 64, 1)]
for prob, H, L in cases:
    for B in (64, 1024, 8192, 65536, 524288):
        torch.manual_seed(0)
        gen = torch.Generator().manual_seed(1)
        z = torch.zeros(B, 1)
        if prob == "heat":
            net = dgm_net.DGM(2, 1, H, L).cuda()
            x = torch.pi * torch.rand([B, 1], generator=gen); t = 3.0 * torch.rand([B, 1], generator=gen)
            a = [v.cuda() for v in (torch.cat([x, t], 1), torch.cat([x, z], 1), torch.cat([z, t], 1), torch.cat([z + torch.pi, t], 1), z, z.clone())]
            fn = lambda: K.heat_step(net.desc, net.flat_theta(), *a)
        elif prob == "ode":
            import io, contextlib
            with contextlib.redirect_stdout(io.StringIO()):
                net = neural_networks.MLP(1, 1, H, L, activation="relu").cuda()
            a = [v.cuda() for v in (1.01 * torch.rand([B, 1], generator=gen), z, 2.0 * torch.ones(B, 1))]
            fn = lambda: K.ode_step(net.desc, net.flat_theta(), *a)
        else:
            net = dgm_net.DGM(1, 2, H, L).cuda()
            a = [v.cuda() for v in (30.01 * torch.rand([B, 1], generator=gen), z, torch.zeros(B, 2))]
            fn = lambda: K.fhn_step(net.desc, net.flat_theta(), *a)
        reps = 200 if B <= 8192 else 10
        r = {}
        for eng in (2, 0):   # 2 = tile forced on, 0 = off
            lib.dgmk_set_tile_engine(eng)
            r[eng] = bench(fn, reps)
        lib.dgmk_set_tile_engine(1)
        print(f"{prob} H={H} L={L} B={B}: tile {r[2]:.3f} ms  layer-wise {r[0]:.3f} ms  ratio {r[0] / r[2]:.2f}", flush=True)
