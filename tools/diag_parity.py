"""Diagnostic: per-tensor error of the CUDA step vs the oracle port in FP32 and FP64 on the GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from test_gpu_large import _ref_chunked, _rows
from conftest import rel
from oracle import ref_port as rp
from differential_equations_dnn_b200 import kernels as K, dgm_net, neural_networks

which = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
torch.manual_seed(1234)
gen = torch.Generator().manual_seed(22)
if which == "fhn_mlp":
    net = neural_networks.MLP(1, 2, 128, 3, activation="tanh").cuda(); spec = rp.NetSpec(rp.KIND_MLP, 1, 2, 128, 3, rp.ACT_TANH)
elif which == "fhn_dgm":
    net = dgm_net.DGM(1, 2, 128, 4).cuda(); spec = rp.NetSpec(rp.KIND_DGM_LINEAR, 1, 2, 128, 4, rp.ACT_TANH)
a = [(30.01 * torch.rand([B, 1], generator=gen)).cuda(), torch.zeros(B, 1).cuda(), torch.zeros(B, 2).cuda()]
lib = __import__("differential_equations_dnn_b200._cabi", fromlist=["x"]).load()
outs = {}
for eng in (1, 0):
    lib.dgmk_set_gemm_engine(eng)
    outs[eng] = K.fhn_step(net.desc, net.flat_theta(), *a).double().cpu().numpy()
lib.dgmk_set_gemm_engine(1)
l32, g32 = _ref_chunked(rp.fhn_loss, spec, net.flat_theta(), a, B, 1 << 16, _rows)
l64, g64 = _ref_chunked(rp.fhn_loss, spec, net.flat_theta(), a, B, 1 << 16, _rows, torch.float64)
print("loss ours(tc) %.9g ours(ffma) %.9g ref32 %.9g ref64 %.9g" % (outs[1][-1], outs[0][-1], l32, l64))
names = [n for n, _ in net.named_parameters()]
for (p, off, n, live), nm in zip(net.param_slices(), names):
    s = slice(off, off + n)
    print("%-24s |g| %.3e  tc-vs-64 %.2e  ffma-vs-64 %.2e  ref32-vs-64 %.2e  tc-vs-32 %.2e" % (
        nm, np.linalg.norm(g64[s]), rel(outs[1][:-1][s], g64[s]), rel(outs[0][:-1][s], g64[s]), rel(g32[s], g64[s]), rel(outs[1][:-1][s], g32[s])))
