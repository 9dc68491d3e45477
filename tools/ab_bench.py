"""Scratch A/B timing of the heat step at the bench workload (not the contract bench):
ms per step and the per-kernel-class times from dgmk_profile.  python tools/ab_bench.py [tag]"""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from differential_equations_dnn_b200 import kernels as K, _cabi


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else ""
    lib = _cabi.load()
    H, L, B = 128, 3, 1 << 20
    d = _cabi.make_desc(_cabi.KIND_DGM_LINEAR, 2, 1, H, L, _cabi.ACT_TANH)
    P = K.param_count(d)
    torch.manual_seed(0)
    theta = ((torch.rand(P) - 0.5) * 0.2).cuda()
    x = torch.pi * torch.rand(B, 1, device="cuda"); t = 3 * torch.rand(B, 1, device="cuda")
    z = torch.zeros(B, 1, device="cuda")
    X, X0, B1, B2 = torch.cat([x, t], 1), torch.cat([x, z], 1), torch.cat([z, t], 1), torch.cat([z + torch.pi, t], 1)
    for _ in range(2):
        out = K.heat_step(d, theta, X, X0, B1, B2, z, z)
    torch.cuda.synchronize()
    n = 4
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = K.heat_step(d, theta, X, X0, B1, B2, z, z)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    lib.dgmk_profile(1)
    for _ in range(2):
        out = K.heat_step(d, theta, X, X0, B1, B2, z, z)
    torch.cuda.synchronize()
    lib.dgmk_profile(0)
    cls = []
    for c in range(5):
        t_, n_, f_, b_ = C.c_double(), C.c_longlong(), C.c_double(), C.c_double()
        if lib.dgmk_profile_read(c, C.byref(t_), C.byref(n_), C.byref(f_), C.byref(b_)) == 0:
            cls.append(f"{c}:{t_.value / 2:.2f}ms/{n_.value // 2}")
    print(f"{tag} {ms:.2f} ms/step {B / ms * 1e3:.3e} rows/s | " + " ".join(cls) + f" | loss {out[-1].item():.6f}", flush=True)


main()
