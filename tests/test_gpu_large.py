"""-m gpu: BASELINE-size runs checked through size-independent properties, plus the
torch-autograd oracle executed on the GPU (SURVEY 8c tier 3) at the largest size it fits."""
import numpy as np
import pytest
import torch

from conftest import grad_groups, rel

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


# ------------------------------------------------------------------------------------------------
# BASELINE-size parity against the torch-autograd oracle executed ON THE GPU (SURVEY 8c tier 3:
# oracle/ref_port = the reference's nested-autograd algorithm), chunk-accumulated so that the
# double-backward graph fits, in FP32 (cuBLAS SGEMM, TF32 off: what the reference computes) AND in FP64
# (the arbiter, SURVEY 8c tier 2).  Bar (north_star): loss and per-tensor norm-wise gradient within 1e-5
#   * of the FP64 oracle, always;
#   * of the FP32 oracle, widened to 2 x the FP32 oracle's own distance from FP64 when that is larger:
#     at 2^20 rows cuBLAS's FP32 accumulation puts the reference itself 1.7e-5 from FP64 on the hidden
#     weights of MLP(1,2,128,3) (measured: tools/diag_parity.py; this path stays within 3.5e-6), and ReLU
#     networks flip a derivative wherever a pre-activation lies within rounding of 0.
def _ref_chunked(loss_fn, spec, theta, args, B, CH, slicer, dtype=torch.float32):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    gsum = torch.zeros_like(theta, dtype=torch.float64)
    lsum = 0.0
    for lo in range(0, B, CH):
        hi = min(B, lo + CH)
        th = theta.detach().to(dtype).clone().requires_grad_(True)
        l = loss_fn(spec, th, *[z.to(dtype) for z in slicer(args, lo, hi)]) * ((hi - lo) / B)
        l.backward()
        gsum += th.grad.double()
        lsum += float(l.item())
    return lsum, gsum.cpu().numpy()


def _rows(args, lo, hi):
    return [z[lo:hi] for z in args]


def _worst(a, b, layout):
    return max([rel(a[off:off + n], b[off:off + n]) for off, n, live in grad_groups(layout)
                if live and np.linalg.norm(b[off:off + n]) > 0] + [0.0])


def _parity(out, loss_fn, spec, net, args, B, CH, slicer=_rows, tag="", tol=1e-5, cap=1e-4):
    layout = [(off, n, live) for _, off, n, live in net.param_slices()]
    out = out.double().cpu().numpy()
    l32, g32 = _ref_chunked(loss_fn, spec, net.flat_theta(), args, B, CH, slicer)
    l64, g64 = _ref_chunked(loss_fn, spec, net.flat_theta(), args, B, CH, slicer, torch.float64)
    e64 = max(_worst(out[:-1], g64, layout), abs(out[-1] - l64) / abs(l64))
    ref_noise = max(_worst(g32, g64, layout), abs(l32 - l64) / abs(l64))
    e32 = max(_worst(out[:-1], g32, layout), abs(out[-1] - l32) / abs(l32))
    print(f"{tag}: vs FP64 oracle {e64:.2e}, vs FP32 oracle {e32:.2e}, FP32 oracle vs FP64 {ref_noise:.2e}")
    names = [nm for nm, _ in net.named_parameters()]
    per = [(rel(out[off:off + n], g64[off:off + n]), rel(g32[off:off + n], g64[off:off + n]), nm, n)
           for (off, n, live), nm in zip(layout, names) if live and np.linalg.norm(g64[off:off + n]) > 0]
    for e, r, nm, n in sorted(per, reverse=True)[:4]:
        print(f"    {nm} [{n}]: ours vs FP64 {e:.2e}, FP32 oracle vs FP64 {r:.2e}")
    for off, n, live in layout:
        if not live:
            assert np.all(out[off:off + n] == 0)
    return e64, e32, ref_noise


def _assert_parity(e64, e32, ref_noise, tol=1e-5, cap=1e-4, strict64=True):
    tol32 = max(tol, 2 * ref_noise)
    assert tol32 <= cap, ref_noise
    assert e32 < tol32, (e32, tol32)
    # strict64 = False only for ReLU networks: FP64 lands on the other side of a tie than any FP32 evaluation,
    # ours or the reference's
    assert e64 < (tol if strict64 else tol32), e64


def test_heat_dgm128_full_2_20_vs_oracle():
    """BASELINE configs[1] at its full size: heat + dgm_net.DGM(2,1,128,3), 2^20 rows, 16 chunks of 65536."""
    from oracle import ref_port as rp
    from differential_equations_dnn_b200 import kernels as K
    net = make(128, 3)
    B = 1 << 20
    a = heat_inputs(B, 21)
    out = K.heat_step(net.desc, net.flat_theta(), *a)
    spec = rp.NetSpec(rp.KIND_DGM_LINEAR, 2, 1, 128, 3, rp.ACT_TANH)
    _assert_parity(*_parity(out, rp.heat_loss, spec, net, a, B, 1 << 16, tag="heat DGM(2,1,128,3) 2^20"))


@pytest.mark.parametrize("kind", ["mlp", "dgm"])
def test_fhn_full_2_20_vs_oracle(kind):
    """BASELINE configs[2]: FitzHugh-Nagumo loss at 2^20 time points, MLP(1,2,128,3,tanh) (the config) and the
    as-shipped dgm_net.DGM(1,2,128,4) (fitzhugh_nagumo.py:53-97, :211-214)."""
    from oracle import ref_port as rp
    from differential_equations_dnn_b200 import kernels as K, dgm_net, neural_networks
    torch.manual_seed(1234)
    if kind == "mlp":
        net = neural_networks.MLP(1, 2, 128, 3, activation="tanh").cuda()
        spec = rp.NetSpec(rp.KIND_MLP, 1, 2, 128, 3, rp.ACT_TANH)
    else:
        net = dgm_net.DGM(1, 2, 128, 4).cuda()
        spec = rp.NetSpec(rp.KIND_DGM_LINEAR, 1, 2, 128, 4, rp.ACT_TANH)
    B = 1 << 20
    gen = torch.Generator().manual_seed(22)
    a = [(30.01 * torch.rand([B, 1], generator=gen)).cuda(), torch.zeros(B, 1).cuda(), torch.zeros(B, 2).cuda()]
    out = K.fhn_step(net.desc, net.flat_theta(), *a)
    _assert_parity(*_parity(out, rp.fhn_loss, spec, net, a, B, 1 << 16, tag="fhn 2^20 " + kind))


@pytest.mark.parametrize("act", ["relu", "tanh"])
def test_ode_full_2_20_vs_oracle(act):
    """BASELINE configs[0] at the throughput size: simple_ode loss, MLP(1,1,32) (relu as shipped,
    simple_ode.py:167; tanh variant), 2^20 rows."""
    from oracle import ref_port as rp
    from differential_equations_dnn_b200 import kernels as K, neural_networks
    torch.manual_seed(1234)
    net = neural_networks.MLP(1, 1, 32, 1, activation=act).cuda()
    spec = rp.NetSpec(rp.KIND_MLP, 1, 1, 32, 1, rp.ACT_NAMES[act])
    B = 1 << 20
    gen = torch.Generator().manual_seed(23)
    a = [(1.01 * torch.rand([B, 1], generator=gen)).cuda(), torch.zeros(B, 1).cuda(), 2.0 * torch.ones(B, 1).cuda()]
    out = K.ode_step(net.desc, net.flat_theta(), *a)
    _assert_parity(*_parity(out, rp.ode_loss, spec, net, a, B, 1 << 18, tag="ode 2^20 " + act), strict64=False)


def test_fredholm_k1024_vs_oracle():
    """BASELINE configs[3]: neural_networks.DGM(1,1,32) (ReLU gates, raw parameters), k = 1024 quadrature nodes
    per point, B = 2^12 points = 4.2 M node evaluations per step (fredholm.py:47-74)."""
    from oracle import ref_port as rp
    from differential_equations_dnn_b200 import kernels as K, neural_networks
    torch.manual_seed(1234)
    net = neural_networks.DGM(1, 1, 32, 1).cuda()
    # biases start at exactly 0 there (neural_networks.py:93-96): the reference init, ReLU ties included
    spec = rp.NetSpec(rp.KIND_DGM_RAW, 1, 1, 32, 1, rp.ACT_RELU)
    B, k = 1 << 12, 1024
    gen = torch.Generator().manual_seed(24)
    x = ((np.pi / 2) * torch.rand(B, 1, generator=gen)).cuda()
    T = ((np.pi / 2) * torch.rand(k, B, 1, generator=gen)).cuda()
    out = K.fredholm_step(net.desc, net.flat_theta(), x, T)

    def sl(args, lo, hi):
        return [args[0][lo:hi], args[1][:, lo:hi]]
    _assert_parity(*_parity(out, rp.fredholm_loss, spec, net, [x, T], B, 1 << 11, sl, tag="fredholm k=1024"), strict64=False)


def test_fhn_shipped_grid_sampler():
    """The shipped FitzHugh-Nagumo sampler (fitzhugh_nagumo.py:123-133: `batch_size` distinct nodes of a
    200-point grid drawn with torch.multinomial) on the GPU: the first iteration's loss equals the oracle's on
    the same draw, and the driver runs (eagerly and replayed from a CUDA graph)."""
    from oracle import ref_port as rp
    from differential_equations_dnn_b200 import dgm_net, fitzhugh_nagumo as fz
    spec = rp.NetSpec(rp.KIND_DGM_LINEAR, 1, 2, 32, 2, rp.ACT_TANH)
    bs = 100   # the CLI default (fitzhugh_nagumo.py:202-204)
    y_ic = torch.zeros(bs, 2, device="cuda")
    for graph in (False, True):
        torch.manual_seed(1234)
        net = dgm_net.DGM(1, 2, 32, 2).cuda()
        theta0 = net.flat_theta().clone()
        torch.manual_seed(77)
        _, losses = fz.minimize_loss_dgm(net, y_ic, iterations=40, batch_size=bs, lrate=1e-3, sampler="grid", cuda_graph=graph)
        assert len(losses) == 40 and np.all(np.isfinite(losses))
        torch.manual_seed(77)   # replay the first draw
        Tg = torch.linspace(0.0, 30.0, steps=200, device="cuda")
        prob = torch.full((200,), 1.0 / 200, device="cuda")
        t = Tg[prob.multinomial(num_samples=bs, replacement=False)].reshape(-1, 1)
        assert t.unique().numel() == bs
        l0 = rp.fhn_loss(spec, theta0, t, torch.zeros(bs, 1, device="cuda"), y_ic)
        assert abs(losses[0] - l0.item()) <= 1e-5 * abs(l0.item()), (graph, losses[0], l0.item())
        assert np.mean(losses[-5:]) < np.mean(losses[:5])
