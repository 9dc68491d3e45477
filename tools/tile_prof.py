"""One fused step of a small-hidden-size config (for ncu captures of tk::tile_step_kernel).
   python tools/tile_prof.py heat|ode|fhn H L rows [reps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from differential_equations_dnn_b200 import dgm_net, neural_networks, kernels as K  # noqa: E402

prob, H, L, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
torch.manual_seed(1234)
gen = torch.Generator().manual_seed(1)
z = torch.zeros(B, 1)
if prob == "heat":
    net = dgm_net.DGM(2, 1, H, L).cuda()
    x = torch.pi * torch.rand([B, 1], generator=gen); t = 3.0 * torch.rand([B, 1], generator=gen)
    a = [v.cuda() for v in (torch.cat([x, t], 1), torch.cat([x, z], 1), torch.cat([z, t], 1), torch.cat([z + torch.pi, t], 1), z, z.clone())]
    fn = lambda: K.heat_step(net.desc, net.flat_theta(), *a)
elif prob == "ode":
    net = neural_networks.MLP(1, 1, H, L, activation="relu").cuda()
    a = [v.cuda() for v in (1.01 * torch.rand([B, 1], generator=gen), z, 2.0 * torch.ones(B, 1))]
    fn = lambda: K.ode_step(net.desc, net.flat_theta(), *a)
elif prob == "fredholm":
    k = int(os.environ.get("K", "1024"))
    net = neural_networks.DGM(1, 1, H, L).cuda()
    a = [((torch.pi / 2) * torch.rand([B, 1], generator=gen)).cuda(), ((torch.pi / 2) * torch.rand([k, B, 1], generator=gen)).cuda()]
    fn = lambda: K.fredholm_step(net.desc, net.flat_theta(), *a)
else:
    net = dgm_net.DGM(1, 2, H, L).cuda()
    a = [v.cuda() for v in (30.01 * torch.rand([B, 1], generator=gen), z, torch.zeros(B, 2))]
    fn = lambda: K.fhn_step(net.desc, net.flat_theta(), *a)
if os.environ.get("TILE_ENGINE"):   # 0: layer-wise only, 2: the resident-tile step wherever it fits
    from differential_equations_dnn_b200 import _cabi as _c
    _c.load().dgmk_set_tile_engine(int(os.environ["TILE_ENGINE"]))
for _ in range(2):
    out = fn()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    out = fn()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
print(f"{prob} H={H} L={L} B={B}: {dt * 1e3:.3f} ms/step, {B / dt:.4g} rows/s, loss {out[-1].item():.6f}")

# stage timeline of CTA 0 (first tiles of the launch)
if os.environ.get("TIMELINE"):
    import ctypes as C
    from differential_equations_dnn_b200 import _cabi
    lib = _cabi.load()
    N = 400
    buf = torch.zeros(2 * N, dtype=torch.int64, device="cuda")
    lib.dgmk_tile_profile(C.c_void_p(buf.data_ptr()), N)
    fn()
    torch.cuda.synchronize()
    lib.dgmk_tile_profile(None, 0)
    b = buf.cpu().numpy().reshape(N, 2)
    kinds = {0: "start", 1: "ew", 2: "ew4", 3: "gemm_nn", 4: "colsum", 5: "gemm_tn", 6: "AtE", 7: "rowdot"}
    n = int((b[:, 0] > 0).sum())
    print("stages recorded", n)
    tot = {}
    prev = b[0, 0]
    line = []
    for i in range(1, n):
        dt = int(b[i, 0] - prev); prev = b[i, 0]
        k = kinds[int(b[i, 1])]
        tot[k] = tot.get(k, 0) + dt
        line.append(f"{k}:{dt}")
    per = int(os.environ.get("STAGES_PER_LINE", "12"))
    for i in range(0, min(len(line), 120), per):
        print("  ", " ".join(line[i:i + per]))
    s = sum(tot.values())
    print("totals (cycles over the recorded stages):", {k: (v, f"{100 * v / s:.1f}%") for k, v in sorted(tot.items(), key=lambda kv: -kv[1])})
