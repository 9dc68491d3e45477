"""Scratch timing of the heat step (not the contract bench)."""
import sys, time
import torch
sys.path.insert(0, ".")
from differential_equations_dnn_b200 import kernels as K, _cabi

def main():
    H, L = 128, 3
    d = _cabi.make_desc(_cabi.KIND_DGM_LINEAR, 2, 1, H, L, _cabi.ACT_TANH)
    P = K.param_count(d)
    torch.manual_seed(0)
    theta = ((torch.rand(P) - 0.5) * 0.2).cuda()
    for logB in (14, 17, 20):
        B = 1 << logB
        x = torch.pi * torch.rand(B, 1, device="cuda"); t = 3 * torch.rand(B, 1, device="cuda")
        z = torch.zeros(B, 1, device="cuda")
        X, X0, B1, B2 = torch.cat([x, t], 1), torch.cat([x, z], 1), torch.cat([z, t], 1), torch.cat([z + torch.pi, t], 1)
        for _ in range(2):
            out = K.heat_step(d, theta, X, X0, B1, B2, z, z)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3
        e0.record()
        for _ in range(n):
            out = K.heat_step(d, theta, X, X0, B1, B2, z, z)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        flops = 8262912.0 * B
        print(f"B=2^{logB}: {ms:.2f} ms/step  {B/ms*1e3:.3e} rows/s  {flops/ms*1e-9:.2f} TFLOP/s alg  loss={out[-1].item():.5f}  ws={K.workspace_bytes(d,0,B)/2**30:.2f} GiB", flush=True)

main()
