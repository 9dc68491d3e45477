"""Per-tensor difference between the tcgen05 3xTF32 engine and the FP32 FFMA2 engine."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from differential_equations_dnn_b200 import kernels as K, _cabi, dgm_net
lib = _cabi.load()
def rel(a, b): return float(np.linalg.norm(a.double().cpu().numpy() - b.double().cpu().numpy()) / max(np.linalg.norm(b.double().cpu().numpy()), 1e-300))
def inputs(B, seed):
    gen = torch.Generator().manual_seed(seed)
    x = torch.pi * torch.rand([B, 1], generator=gen); t = 3.0 * torch.rand([B, 1], generator=gen); z = torch.zeros(B, 1)
    return [a.cuda() for a in (torch.cat([x, t], 1), torch.cat([x, z], 1), torch.cat([z, t], 1), torch.cat([z + torch.pi, t], 1), z, z.clone())]
for (H, L, B) in ((128, 3, 1 << 16), (128, 1, 4096)):
    torch.manual_seed(1234); net = dgm_net.DGM(2, 1, H, L).cuda()
    a = inputs(B, 5)
    lib.dgmk_set_gemm_engine(0); ref = K.heat_step(net.desc, net.flat_theta(), *a).clone()
    lib.dgmk_set_gemm_engine(1); tc = K.heat_step(net.desc, net.flat_theta(), *a).clone()
    print(f"H={H} L={L} B={B}: loss ffma {ref[-1].item():.8f} tc {tc[-1].item():.8f} rel {abs(ref[-1].item()-tc[-1].item())/abs(ref[-1].item()):.2e}")
    worst = []
    for (name, p), (pp, off, n, live) in zip(net.named_parameters(), net.param_slices()):
        worst.append((rel(tc[off:off+n], ref[off:off+n]), name))
    worst.sort(reverse=True)
    print("  worst tensors:", [(f"{e:.2e}", n) for e, n in worst[:6]])
    print("  whole grad:", f"{rel(tc[:-1], ref[:-1]):.2e}")
