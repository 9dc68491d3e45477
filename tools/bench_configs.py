"""Throughput of the other BASELINE.json configs / the SURVEY 8(d) size sweep (fused step only,
inputs resident, CUDA events, 3 warm-up + 5 timed).  Not the contract bench (bench.py)."""
import sys, json
import torch
sys.path.insert(0, ".")
from differential_equations_dnn_b200 import kernels as K, _cabi

def f_alg(kind, H, L, o, M):
    c = 2 if kind == 0 else 8
    return 3 * M * (L * c * H * H + 2 * H * o)

def timeit(fn, warm=3, n=5):
    for _ in range(warm): out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out

def theta_for(d):
    torch.manual_seed(0)
    return ((torch.rand(K.param_count(d)) - 0.5) * 0.2).cuda()

rows = []
def report(name, B, ms, flops_row, extra=""):
    r = dict(config=name, rows=B, ms_per_step=round(ms, 3), rows_per_s=B / ms * 1e3, alg_tflops=flops_row * B / ms * 1e-9)
    rows.append(r)
    print(f"{name:58s} B={B:>8d}  {ms:9.3f} ms  {B/ms*1e3:10.3e} rows/s  {flops_row*B/ms*1e-9:7.2f} TFLOP/s alg {extra}", flush=True)

dev = "cuda"
B = 1 << 20
# C2 + sweep: heat
x = torch.pi * torch.rand(B, 1, device=dev); t = 3 * torch.rand(B, 1, device=dev); z = torch.zeros(B, 1, device=dev)
X, X0, B1, B2 = torch.cat([x, t], 1), torch.cat([x, z], 1), torch.cat([z, t], 1), torch.cat([z + torch.pi, t], 1)
for kind, H, L, act, nm in ((1, 128, 3, 2, "heat dgm_net.DGM(2,1,128,3)"), (1, 128, 4, 2, "heat dgm_net.DGM(2,1,128,4)"),
                            (1, 64, 3, 2, "heat dgm_net.DGM(2,1,64,3)"), (1, 50, 3, 2, "heat dgm_net.DGM(2,1,50,3)"),
                            (1, 32, 1, 2, "heat dgm_net.DGM(2,1,32,1)"), (0, 128, 3, 2, "heat MLP(2,1,128,3,tanh)"),
                            (0, 128, 3, 0, "heat MLP(2,1,128,3,relu) [as shipped]")):
    d = _cabi.make_desc(kind, 2, 1, H, L, act); th = theta_for(d)
    ms, _ = timeit(lambda: K.heat_step(d, th, X, X0, B1, B2, z, z))
    report(nm, B, ms, f_alg(kind, H, L, 1, 7))
# C1 simple_ode
tt = 1.01 * torch.rand(B, 1, device=dev); yic = torch.full((B, 1), 2.0, device=dev)
for act, nm in ((0, "simple_ode MLP(1,1,32,1,relu)"), (2, "simple_ode MLP(1,1,32,1,tanh)")):
    d = _cabi.make_desc(0, 1, 1, 32, 1, act); th = theta_for(d)
    ms, _ = timeit(lambda: K.ode_step(d, th, tt, z, yic))
    report(nm, B, ms, f_alg(0, 32, 1, 1, 3))
# C3 FHN
t30 = 30.01 * torch.rand(B, 1, device=dev); y2 = torch.zeros(B, 2, device=dev)
for kind, H, L, nm in ((0, 128, 3, "fhn MLP(1,2,128,3,tanh)"), (1, 128, 4, "fhn dgm_net.DGM(1,2,128,4) [as shipped]")):
    d = _cabi.make_desc(kind, 1, 2, H, L, 2); th = theta_for(d)
    ms, _ = timeit(lambda: K.fhn_step(d, th, t30, z, y2))
    report(nm, B, ms, f_alg(kind, H, L, 2, 3))
# C4 Fredholm
Bf, k = 1 << 14, 1024
xf = (torch.pi / 2) * torch.rand(Bf, 1, device=dev); T = (torch.pi / 2) * torch.rand(k, Bf, 1, device=dev)
d = _cabi.make_desc(2, 1, 1, 32, 1, 0); th = theta_for(d)
ms, _ = timeit(lambda: K.fredholm_step(d, th, xf, T), warm=2, n=3)
report("fredholm neural_networks.DGM(1,1,32,1), k=1024", Bf, ms, f_alg(2, 32, 1, 1, k + 1), f"({Bf*(k+1)/ms*1e3:.3e} node-evals/s)")
json.dump(rows, open("gpurun_out/configs.json", "w"), indent=1)
