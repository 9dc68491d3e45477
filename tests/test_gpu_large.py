"""-m gpu: BASELINE-size runs checked through size-independent properties, plus the
torch-autograd oracle executed on the GPU (SURVEY 8c tier 3) at the largest size it fits."""
import numpy as np
import pytest
import torch

from conftest import rel

pytestmark = pytest.mark.gpu


def heat_inputs(B, seed, dev="cuda"):
    gen = torch.Generator().manual_seed(seed)
    x = torch.pi * torch.rand([B, 1], generator=gen)
    t = 3.0 * torch.rand([B, 1], generator=gen)
    z = torch.zeros(B, 1)
    return [a.to(dev) for a in (torch.cat([x, t], 1), torch.cat([x, z], 1), torch.cat([z, t], 1),
                                torch.cat([z + torch.pi, t], 1), z, z.clone())]


def make(H, L, seed=1234):
    from differential_equations_dnn_b200 import dgm_net
    torch.manual_seed(seed)
    return dgm_net.DGM(2, 1, H, L).cuda()


def test_vs_torch_autograd_on_gpu_65536():
    """heat + DGM(2,1,128,3), B = 65536: our kernels vs the oracle port run by torch on
    the same GPU (FP32 cuBLAS, TF32 off)."""
    from oracle import ref_port as rp
    from differential_equations_dnn_b200 import kernels as K
    torch.backends.cuda.matmul.allow_tf32 = False
    net = make(128, 3)
    a = heat_inputs(1 << 16, 5)
    out = K.heat_step(net.desc, net.flat_theta(), *a)
    spec = rp.NetSpec(rp.KIND_DGM_LINEAR, 2, 1, 128, 3, rp.ACT_TANH)
    gsum = torch.zeros_like(net.flat_theta(), dtype=torch.float64)
    lsum = 0.0
    CH = 1 << 14
    for lo in range(0, 1 << 16, CH):   # chunk-accumulated, mean over the FULL batch
        th = net.flat_theta().detach().clone().requires_grad_(True)
        l = rp.heat_loss(spec, th, *[z[lo:lo + CH] for z in a]) * (CH / (1 << 16))
        l.backward()
        gsum += th.grad.double()
        lsum += l.item()
    assert abs(out[-1].item() - lsum) <= 1e-5 * abs(lsum)
    P = net.flat_theta().numel()
    for p, off, n, live in net.param_slices():
        assert rel(out[off:off + n].cpu().numpy(), gsum[off:off + n].cpu().numpy()) < 1e-5


def test_full_size_properties_2_20():
    """B = 2^20 (BASELINE configs[1]): (a) chunking invariance, (b) shard additivity,
    (c) mean-of-means against 16 sub-batches."""
    from differential_equations_dnn_b200 import kernels as K, _cabi
    net = make(128, 3)
    B = 1 << 20
    a = heat_inputs(B, 7)
    th = net.flat_theta()
    full = K.heat_step(net.desc, th, *a).clone()
    assert torch.isfinite(full).all()
    small = torch.empty(K.workspace_bytes(net.desc, _cabi.WS_HEAT, 50_000), dtype=torch.uint8, device="cuda")
    alt = K.heat_step(net.desc, th, *a, ws=small)
    assert rel(alt.cpu().numpy(), full.cpu().numpy()) < 1e-5
    del small
    acc = torch.zeros_like(full, dtype=torch.float64)
    S = 16
    for s in range(S):
        sl = slice(s * B // S, (s + 1) * B // S)
        acc += K.heat_step(net.desc, th, *[z[sl] for z in a], B_global=B).double()
    assert rel(acc.cpu().numpy(), full.cpu().numpy()) < 1e-5


def test_zero_residual_known_answer():
    """KAT: for u = sin(x) exp(-t) the heat residual is 0.  A net cannot represent it
    exactly, so check the loss kernel's algebra instead: with kappa = 0 the interior term
    reduces to mean(u_t^2), which the order-1 jets reproduce independently."""
    from differential_equations_dnn_b200 import kernels as K, _cabi
    import ctypes as C
    net = make(32, 1)
    B = 4096
    a = heat_inputs(B, 9)
    th = net.flat_theta()
    ws = K.get_workspace(th.device, K.workspace_bytes(net.desc, _cabi.WS_HEAT, B))
    out = torch.empty(th.numel() + 1, device="cuda")
    lib = _cabi.load()
    rc = lib.dgmk_heat_step(C.byref(net.desc), *[C.c_void_p(t.data_ptr()) for t in (th, *a)], B, B, 0.0,
                            C.c_void_p(out.data_ptr() + 4 * th.numel()), C.c_void_p(out.data_ptr()),
                            C.c_void_p(ws.data_ptr()), ws.numel(), None)
    assert rc == 0
    torch.cuda.synchronize()
    Y, J, _, _ = K.jet_forward(net.desc, th, a[0], 1)
    y0 = K.evaluate(net.desc, th, a[1])
    yb1, yb2 = K.evaluate(net.desc, th, a[2]), K.evaluate(net.desc, th, a[3])
    expect = (J[:, 0, 1] ** 2).mean() + ((y0[:, 0] - torch.sin(a[1][:, 0])) ** 2).mean() \
        + (yb1 ** 2).mean() + (yb2 ** 2).mean()
    assert abs(out[-1].item() - expect.item()) <= 1e-5 * abs(expect.item())


def test_fredholm_k1024_tall_operands():
    """BASELINE configs[3]: 2^14 points x 1024 quadrature nodes = 16.8 M value rows in one chunk --
    taller than one gridDim.y worth of row tiles (the GEMM front end goes in slabs).  Property:
    the step of the whole batch equals the sum of the steps of two unequal shards (same B_global)."""
    from differential_equations_dnn_b200 import kernels as K, _cabi
    torch.manual_seed(0)
    d = _cabi.make_desc(_cabi.KIND_DGM_RAW, 1, 1, 32, 1, _cabi.ACT_RELU)
    theta = ((torch.rand(K.param_count(d)) - 0.5) * 0.4).cuda()
    B, k = 1 << 14, 1024
    gen = torch.Generator().manual_seed(4)
    x = ((torch.pi / 2) * torch.rand(B, 1, generator=gen)).cuda()
    T = ((torch.pi / 2) * torch.rand(k, B, 1, generator=gen)).cuda()
    whole = K.fredholm_step(d, theta, x, T).double().cpu().numpy()
    assert np.all(np.isfinite(whole))
    cut = 5000
    parts = 0.0
    for lo, hi in ((0, cut), (cut, B)):
        parts = parts + K.fredholm_step(d, theta, x[lo:hi].contiguous(), T[:, lo:hi].contiguous(), B_global=B).double().cpu().numpy()
    assert abs(parts[-1] - whole[-1]) <= 1e-5 * abs(whole[-1])
    assert np.linalg.norm(parts[:-1] - whole[:-1]) <= 1e-5 * np.linalg.norm(whole[:-1])
