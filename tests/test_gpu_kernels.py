"""-m gpu: the CUDA path (through the C ABI) against the executed-reference golden
vectors and the oracle.  Tolerance: 1e-5 relative on the loss and per-tensor norm-wise
on the gradient (BASELINE.json north_star; SURVEY 7.3 H9 explains norm-wise)."""
import numpy as np
import pytest
import torch

from conftest import golden, golden_names, grad_groups, rel

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def K():
    from differential_equations_dnn_b200 import kernels, _cabi
    _cabi.load()
    return kernels


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def desc_of(g):
    from differential_equations_dnn_b200 import _cabi
    return _cabi.make_desc(*[int(v) for v in g["spec"]])


def run(K, prob, g, ws=None, Bg=None):
    d = desc_of(g)
    th = dev(g["theta"])
    if prob == "heat":
        out = K.heat_step(d, th, *[dev(g[n]) for n in ("X", "X0", "XBD1", "XBD2", "x_bd1", "x_bd2")],
                          B_global=Bg, ws=ws)
    elif prob == "ode":
        out = K.ode_step(d, th, dev(g["t"]), dev(g["t0"]), dev(g["y_ic"]), B_global=Bg, ws=ws)
    elif prob == "fhn":
        out = K.fhn_step(d, th, dev(g["t"]), dev(g["t0"]), dev(g["y_ic"]), B_global=Bg, ws=ws)
    else:
        out = K.fredholm_step(d, th, dev(g["x"]), dev(g["T"]), B_global=Bg, ws=ws)
    out = out.cpu().numpy()
    return float(out[-1]), out[:-1]


def check(K, g, loss, grad, tol=TOL):
    assert abs(loss - float(g["loss"])) <= tol * abs(float(g["loss"])), (loss, float(g["loss"]))
    worst = 0.0
    for off, n, live in grad_groups([(off, r * max(c, 1), live) for off, r, c, live in K.param_layout(desc_of(g))]):
        ref, mine = g["grad"][off:off + n], grad[off:off + n]
        assert np.all(np.isfinite(mine))
        if not live:
            assert np.all(mine == 0)
        elif np.linalg.norm(ref) == 0:
            assert np.linalg.norm(mine) < 1e-7
        else:
            worst = max(worst, rel(mine, ref))
    assert worst < tol, worst


PROBS = ("heat", "ode", "fhn", "fredholm")
CASES = [(p, n) for p in PROBS for n in golden_names(p + "_") if "driver" not in n]


@pytest.fixture
def tile_engine():
    """dgmk_set_tile_engine for the duration of a test (0: layer-wise path only, 1: default dispatch, 2: the resident-tile
    step wherever it fits)."""
    from differential_equations_dnn_b200 import _cabi
    lib = _cabi.load()

    def set_engine(e):
        lib.dgmk_set_tile_engine(e)
    yield set_engine
    lib.dgmk_set_tile_engine(1)


@pytest.mark.parametrize("engine", [1, 0])
@pytest.mark.parametrize("prob,name", CASES)
def test_step_matches_reference(K, tile_engine, prob, name, engine):
    """Every golden fixture through the default dispatch (hidden sizes <= 64: the resident-tile step) and through the
    layer-wise path alone (FFMA2 tiles for the small networks; hidden size 128 runs the tcgen05 kernels either way)."""
    tile_engine(engine)
    g = golden(name)
    loss, grad = run(K, prob, g)
    check(K, g, loss, grad)


@pytest.mark.parametrize("prob,name", [("heat", "heat_dgm_h32l1"), ("fredholm", "fredholm_dgmraw_h32l1_k7"),
                                       ("fhn", "fhn_dgm_h64l2"), ("ode", "ode_mlp_relu_h32l1"),
                                       # hidden size 128: the fused tcgen05 kernels on 5-row chunks (tiles mostly padding)
                                       ("heat", "heat_dgm_h128l3"), ("fhn", "fhn_mlp_tanh_h128l3"), ("fhn", "fhn_dgm_h128l4")])
def test_small_workspace_chunks(K, tile_engine, prob, name):
    from differential_equations_dnn_b200 import _cabi
    g = golden(name)
    if "h128" in name:
        tile_engine(0)   # the chunk loop of the layer-wise path is what this case is about
    cls = {"heat": _cabi.WS_HEAT, "fhn": _cabi.WS_FHN, "fredholm": _cabi.WS_FREDHOLM, "ode": _cabi.WS_ODE}[prob]
    k = g["T"].shape[0] if prob == "fredholm" else 0
    small = K.workspace_bytes(desc_of(g), cls, 5, k)
    ws = torch.empty(small, dtype=torch.uint8, device="cuda")
    loss, grad = run(K, prob, g, ws=ws)
    check(K, g, loss, grad)


def test_dp_shards_sum(K):
    g = golden("heat_dgm_h32l1")
    B = g["X"].shape[0]
    tl, tg = 0.0, 0.0
    for lo, hi in ((0, 23), (23, B)):
        sub = dict(g)
        for n in ("X", "X0", "XBD1", "XBD2", "x_bd1", "x_bd2"):
            sub[n] = g[n][lo:hi]
        l, gr = run(K, "heat", sub, Bg=B)
        tl, tg = tl + l, tg + gr
    check(K, g, tl, tg)


@pytest.mark.parametrize("name", golden_names("jets_"))
def test_jets_and_eval(K, name):
    g = golden(name)
    d = desc_of(g)
    th, X = dev(g["theta"]), dev(g["X"])
    for order in (0, 1, 2):
        Y, J, Hs, ws = K.jet_forward(d, th, X, order)
        assert rel(Y.cpu().numpy(), g["y"]) < TOL
        if order >= 1:
            assert rel(J.cpu().numpy(), g["J"]) < TOL
        if order >= 2:
            assert rel(Hs.cpu().numpy(), g["Hs"]) < TOL   # the FP32 reference itself: 1.4e-7 .. 2.3e-7 vs these FP64 values
    assert rel(K.evaluate(d, th, X).cpu().numpy(), g["y"]) < TOL


@pytest.mark.parametrize("name", golden_names("jets_"))
def test_jet_reverse_vs_oracle(K, name):
    """Random cotangents on (Y, J, Hs): compare with the oracle's autograd."""
    from oracle import ref_port as rp
    g = golden(name)
    d = desc_of(g)
    spec = rp.NetSpec(*[int(v) for v in g["spec"]])
    th = torch.from_numpy(g["theta"]).double().requires_grad_(True)
    X = torch.from_numpy(g["X"]).double().requires_grad_(True)
    y = rp.net_forward(spec, th, X)
    B, o, dd = X.shape[0], spec.o, spec.d
    gen = torch.Generator().manual_seed(11)
    gY = torch.randn(B, o, generator=gen).double()
    gJ = torch.randn(B, o, dd, generator=gen).double()
    gH = torch.randn(B, o, dd, dd, generator=gen).double()
    tot = (y * gY).sum()
    for m in range(o):
        gr = torch.autograd.grad(y[:, m].sum(), X, create_graph=True)[0]
        tot = tot + (gr * gJ[:, m]).sum()
        for i in range(dd):
            hrow = torch.autograd.grad(gr[:, i].sum(), X, create_graph=True)[0]
            tot = tot + (hrow * gH[:, m, i]).sum()
    ref = torch.autograd.grad(tot, th)[0].numpy()
    thd, Xd = dev(g["theta"]), dev(g["X"])
    Y, J, Hs, ws = K.jet_forward(d, thd, Xd, 2)
    grad = K.jet_reverse(d, thd, Xd, 2, dev(gY.numpy()), dev(gJ.numpy()), dev(gH.numpy()), ws).cpu().numpy()
    for off, r, c, live in K.param_layout(d):
        n = r * max(c, 1)
        if np.linalg.norm(ref[off:off + n]) > 0:
            assert rel(grad[off:off + n], ref[off:off + n]) < TOL


def test_adam(K):
    g = golden("adam_5steps")
    th = dev(g["thetas"][0])
    m, v = torch.zeros_like(th), torch.zeros_like(th)
    live = torch.from_numpy(g["live"]).cuda()
    for s in range(5):
        K.adam_step(th, m, v, dev(g["grads"][s]), live, float(g["lr"]), 0.9, 0.999, 1e-8, s + 1)
        assert rel(th.cpu().numpy(), g["thetas"][s + 1]) < 1e-6
    assert rel(m.cpu().numpy(), g["m"]) < 1e-6 and rel(v.cpu().numpy(), g["v"]) < 1e-6
    assert np.array_equal(th.cpu().numpy()[-5:], g["thetas"][0][-5:])  # dead params untouched


def test_host_pointer_is_rejected(K):
    import ctypes as C
    from differential_equations_dnn_b200 import _cabi
    lib = _cabi.load()
    g = golden("ode_mlp_relu_h32l1")
    d = desc_of(g)
    th = np.ascontiguousarray(g["theta"], np.float32)
    p = th.ctypes.data_as(C.c_void_p)
    rc = lib.dgmk_ode_step(C.byref(d), p, p, p, p, 4, 4, p, p, p, 1 << 20, None)
    assert rc == -4 and b"no CPU path" in lib.dgmk_last_error()
    with pytest.raises(Exception):
        K.ode_step(d, torch.from_numpy(th), torch.zeros(4, 1), torch.zeros(4, 1), torch.zeros(4, 1))


def test_vs_oracle_midsize(K):
    """B = 4096 heat + dgm_net.DGM(2,1,64,2): CUDA vs the torch-autograd oracle port."""
    from oracle import ref_port as rp
    from differential_equations_dnn_b200 import _cabi
    spec = rp.NetSpec(rp.KIND_DGM_LINEAR, 2, 1, 64, 2, rp.ACT_TANH)
    gen = torch.Generator().manual_seed(3)
    theta = (torch.rand(spec.num_params(), generator=gen) - 0.5) * 0.3
    B = 4096
    x = torch.pi * torch.rand(B, 1, generator=gen)
    t = 3 * torch.rand(B, 1, generator=gen)
    z = torch.zeros(B, 1)
    X, X0, B1, B2 = torch.cat([x, t], 1), torch.cat([x, z], 1), torch.cat([z, t], 1), torch.cat([z + torch.pi, t], 1)
    loss_ref, g_ref = rp.loss_and_grad(rp.heat_loss, spec, theta, X, X0, B1, B2, z, z)
    d = _cabi.make_desc(spec.kind, spec.d, spec.o, spec.H, spec.L, spec.act)
    out = K.heat_step(d, theta.cuda(), X.cuda(), X0.cuda(), B1.cuda(), B2.cuda(), z.cuda(), z.cuda()).cpu().numpy()
    assert abs(out[-1] - float(loss_ref)) <= TOL * abs(float(loss_ref))
    for off, r, c, live in K.param_layout(d):
        n = r * max(c, 1)
        assert rel(out[off:off + n], g_ref.numpy()[off:off + n]) < TOL


@pytest.mark.parametrize("kind,prob", [("dgm", "heat"), ("mlp", "heat"), ("dgm", "fhn"), ("mlp", "ode128"),
                                       ("dgmraw", "heat"), ("dgmraw", "fredholm")])
def test_engines_agree(K, kind, prob):
    """The fused units-on-lanes kernels + warp-specialised weight gradient (engine 1), the streaming
    tcgen05 tiles + separate element-wise kernels (engine 2) and the FP32 FFMA2 tiles (engine 0)
    compute the same step: each within the parity bar of the FP64 jet oracle, per tensor, on ragged
    row counts (not a multiple of the 64-row tile) at hidden size 128."""
    from differential_equations_dnn_b200 import _cabi, dgm_net, neural_networks
    from oracle import jets_np
    lib = _cabi.load()
    torch.manual_seed(7)
    B = 3000 + 37
    gen = torch.Generator().manual_seed(11)
    if prob == "fredholm":   # neural_networks.DGM (raw [in,out] parameters, ReLU gates) on the value-only channel set
        net = neural_networks.DGM(1, 1, 128, 2).cuda()
        with torch.no_grad():   # biases start at zero there (neural_networks.py:93-96): move off the ReLU ties
            net.flat_theta().add_(0.01 * torch.randn(net.flat_theta().shape, generator=gen).cuda())
        B = 500 + 13
        host = [(np.pi / 2) * torch.rand([B, 1], generator=gen), (np.pi / 2) * torch.rand([6, B, 1], generator=gen)]
        fn, ofn = K.fredholm_step, jets_np.fredholm_step
    elif prob == "heat":
        if kind == "dgmraw":
            net = neural_networks.DGM(2, 1, 128, 2).cuda()
            with torch.no_grad():
                net.flat_theta().add_(0.01 * torch.randn(net.flat_theta().shape, generator=gen).cuda())
        else:
            net = (dgm_net.DGM(2, 1, 128, 2) if kind == "dgm" else neural_networks.MLP(2, 1, 128, 2, activation="tanh")).cuda()
        x = torch.pi * torch.rand([B, 1], generator=gen); t = 3.0 * torch.rand([B, 1], generator=gen); z = torch.zeros(B, 1)
        host = [torch.cat([x, t], 1), torch.cat([x, z], 1), torch.cat([z, t], 1), torch.cat([z + torch.pi, t], 1), z, z.clone()]
        fn, ofn = K.heat_step, jets_np.heat_step
    elif prob == "fhn":
        net = dgm_net.DGM(1, 2, 128, 2).cuda()
        host = [30.01 * torch.rand([B, 1], generator=gen), torch.zeros(B, 1), torch.zeros(B, 2)]
        fn, ofn = K.fhn_step, jets_np.fhn_step
    else:
        net = neural_networks.MLP(1, 1, 128, 2, activation="sigmoid").cuda()
        host = [1.01 * torch.rand([B, 1], generator=gen), torch.zeros(B, 1), 2.0 * torch.ones(B, 1)]
        fn, ofn = K.ode_step, jets_np.ode_step
    d = net.desc
    spec = np.array([d.kind, d.input_dim, d.output_dim, d.hidden_size, d.num_layers, d.activation])
    if kind == "dgmraw":
        # ReLU gates (neural_networks.DGM): a pre-activation within FP32 rounding of 0 lands on the other side of
        # the kink in FP32 than in FP64 and flips that row's derivative -- for ANY FP32 evaluation, the
        # reference's included, and which rows flip depends on the summation order.  So that the 1e-5 bar
        # measures arithmetic and not kink lottery, batch rows with a pre-activation closer than 5e-6 to 0 (FP64
        # oracle; FP32 rounding of these sums is ~1e-7, 3xTF32 ~4e-7) are dropped from the batch (~1 % of rows).
        npnet = jets_np._net(spec, net.flat_theta().double().cpu().numpy())

        def margin(X):
            _, st = jets_np.forward(npnet, jets_np.CS_V, X.double().numpy())
            m = np.abs(st["in"][0][0]).min(1)
            for L in st["layers"]:
                for nm in ("aZ", "aG", "aR", "aH"):
                    m = np.minimum(m, np.abs(L[nm][0]).min(1))
            return m
        if prob == "fredholm":
            m = np.minimum(margin(host[0]), margin(host[1].reshape(-1, 1)).reshape(host[1].shape[0], -1).min(0))
            keep = torch.from_numpy(m > 5e-6)
            host = [host[0][keep], host[1][:, keep]]
        else:
            m = np.minimum.reduce([margin(h) for h in host[:4]])
            keep = torch.from_numpy(m > 5e-6)
            host = [h[keep] for h in host]
        assert keep.sum().item() > 0.9 * keep.numel(), keep.sum().item()
    args = [a.contiguous().cuda() for a in host]
    lo, go = ofn(spec, net.flat_theta().double().cpu().numpy(), *[a.double().numpy() for a in host])
    try:
        out = {}
        for eng in (0, 2, 1):
            lib.dgmk_set_gemm_engine(eng)
            out[eng] = fn(d, net.flat_theta(), *args).double().cpu().numpy()
    finally:
        lib.dgmk_set_gemm_engine(1)
    tol = TOL
    for eng in (0, 2, 1):
        assert abs(out[eng][-1] - lo) <= tol * abs(lo), (eng, out[eng][-1], lo)
        for off, n, live in grad_groups([(off, n, live) for _, off, n, live in net.param_slices()]):
            if live and np.linalg.norm(go[off:off + n]) > 0:
                assert rel(out[eng][off:off + n], go[off:off + n]) < tol, (eng, off, rel(out[eng][off:off + n], go[off:off + n]))


# ------------------------------------------------------------------------------------------------
# The resident-tile step (csrc/dgmk_tile.cuh: one persistent kernel per step for hidden sizes <= 64) against the
# layer-wise path and the FP64 oracle: many tiles per CTA, ragged last tile, several FP32 accumulation segments,
# weights / accumulators in shared memory (hidden size 32) and in L2 (hidden size 64, 3 layers), one and two CTAs per SM.
@pytest.mark.parametrize("case", ["heat_dgm32", "heat_dgm64x3", "heat_mlp64", "ode_mlp32", "fhn_dgm32", "fredholm_dgmraw32", "fredholm_k70",
                                  "heat_dgm128x3", "heat_mlp128x3", "fhn_dgm128x4"])
def test_tile_step_vs_layerwise_and_oracle(K, case):
    from differential_equations_dnn_b200 import _cabi, dgm_net, neural_networks
    from oracle import jets_np
    lib = _cabi.load()
    torch.manual_seed(3)
    gen = torch.Generator().manual_seed(4)
    if case.startswith("heat"):
        B = 203 if "128" in case else 5000 + 13   # hidden size 128: the tile step covers the small-batch regime only
        net = {"heat_dgm32": lambda: dgm_net.DGM(2, 1, 32, 1), "heat_dgm64x3": lambda: dgm_net.DGM(2, 1, 64, 3),
               "heat_mlp64": lambda: neural_networks.MLP(2, 1, 50, 2, activation="sigmoid"),
               "heat_dgm128x3": lambda: dgm_net.DGM(2, 1, 128, 3),
               "heat_mlp128x3": lambda: neural_networks.MLP(2, 1, 128, 3, activation="tanh")}[case]().cuda()
        x = torch.pi * torch.rand([B, 1], generator=gen); t = 3.0 * torch.rand([B, 1], generator=gen); z = torch.zeros(B, 1)
        host = [torch.cat([x, t], 1), torch.cat([x, z], 1), torch.cat([z, t], 1), torch.cat([z + torch.pi, t], 1), z, z.clone()]
        fn, ofn = K.heat_step, jets_np.heat_step
    elif case == "ode_mlp32":
        B = 40000 + 7
        net = neural_networks.MLP(1, 1, 32, 1, activation="tanh").cuda()
        host = [1.01 * torch.rand([B, 1], generator=gen), torch.zeros(B, 1), 2.0 * torch.ones(B, 1)]
        fn, ofn = K.ode_step, jets_np.ode_step
    elif case.startswith("fhn_dgm"):
        B, net = (100, dgm_net.DGM(1, 2, 128, 4).cuda()) if case == "fhn_dgm128x4" else (9000 + 1, dgm_net.DGM(1, 2, 32, 2).cuda())
        host = [30.01 * torch.rand([B, 1], generator=gen), torch.zeros(B, 1), torch.zeros(B, 2)]
        fn, ofn = K.fhn_step, jets_np.fhn_step
    else:
        B, k = (300 + 5, 50) if case == "fredholm_dgmraw32" else (37, 70)
        net = neural_networks.DGM(1, 1, 32, 1).cuda()
        with torch.no_grad():   # off the exact ReLU ties of the zero-initialised biases
            net.flat_theta().add_(0.01 * torch.randn(net.flat_theta().shape, generator=gen).cuda())
        host = [(np.pi / 2) * torch.rand([B, 1], generator=gen), (np.pi / 2) * torch.rand([k, B, 1], generator=gen)]
        fn, ofn = K.fredholm_step, jets_np.fredholm_step
    args = [a.cuda() for a in host]
    d = net.desc
    spec = np.array([d.kind, d.input_dim, d.output_dim, d.hidden_size, d.num_layers, d.activation])
    lo, go = ofn(spec, net.flat_theta().double().cpu().numpy(), *[a.double().numpy() for a in host])
    out = {}
    try:
        for tag, eng, flush in (("tile", 2, 0), ("tile_segments", 2, 2), ("layerwise", 0, 0)):
            lib.dgmk_set_tile_engine(eng)
            lib.dgmk_set_tile_flush(flush)
            n0 = lib.dgmk_launch_count()
            out[tag] = fn(d, net.flat_theta(), *args).double().cpu().numpy()
            launches = lib.dgmk_launch_count() - n0
            assert (launches <= 5) == tag.startswith("tile"), (tag, launches)   # pack + tile kernel + reduce + unpack
    finally:
        lib.dgmk_set_tile_engine(1)
        lib.dgmk_set_tile_flush(0)
    layout = grad_groups([(off, n, live) for _, off, n, live in net.param_slices()])
    relu = "dgmraw" in case or case == "fredholm_k70"
    tol = 3e-5 if relu else TOL   # ReLU: rows within rounding of a kink flip in any FP32 evaluation (see test_engines_agree)
    for tag, o in out.items():
        assert abs(o[-1] - lo) <= tol * abs(lo), (tag, o[-1], lo)
        for off, n, live in layout:
            if live and np.linalg.norm(go[off:off + n]) > 0:
                assert rel(o[off:off + n], go[off:off + n]) < tol, (tag, off, rel(o[off:off + n], go[off:off + n]))
    # the tile step is deterministic: same launch twice -> same bits
    lib.dgmk_set_tile_engine(2)
    try:
        again = fn(d, net.flat_theta(), *args).double().cpu().numpy()
    finally:
        lib.dgmk_set_tile_engine(1)
    assert np.array_equal(again, out["tile"])


# ------------------------------------------------------------------------------------------------
# On-device collocation sampler: the CUDA kernels against oracle/philox_np.py, bit for bit.
@pytest.mark.parametrize("n,lo,hi,seed,stream,step_dev,step_add", [
    (1, 0.0, 1.0, 0, 0, None, 0), (7, 0.0, 1.01, 1234, 0, None, 5), (64, -2.0, 3.5, 2 ** 63 + 12345, 9, 41, 1),
    (100003, 0.0, np.pi / 2, 2 ** 40 + 3, 257, 2 ** 33 + 5, 0), (50 * 32, 0.0, np.pi / 2, 7, 1, 3, 0)])
def test_philox_uniform_kernel_is_bit_exact(K, n, lo, hi, seed, stream, step_dev, step_add):
    from oracle import philox_np as PH
    out = torch.full([n + 5], float("nan"), device="cuda")
    step = None if step_dev is None else torch.tensor([step_dev], dtype=torch.int64, device="cuda")
    K.sample_uniform(out[:n], lo, hi, seed, stream, step, step_add)
    got = out.cpu().numpy()
    assert np.array_equal(got[:n], PH.uniform(n, lo, hi, seed, stream, (step_dev or 0) + step_add))
    assert np.all(np.isnan(got[n:]))


@pytest.mark.parametrize("B", [1, 6, 64, 4099])
def test_philox_heat_kernel_is_bit_exact(K, B):
    from oracle import philox_np as PH
    bufs = [torch.full([B, 2], float("nan"), device="cuda") for _ in range(4)]
    step = torch.tensor([17], dtype=torch.int64, device="cuda")
    K.sample_heat(*bufs, np.pi, 3.0, np.pi, 99, step, 2)
    for got, want in zip(bufs, PH.heat(B, np.pi, 3.0, np.pi, 99, 19)):
        assert np.array_equal(got.cpu().numpy(), want)
    # a captured graph draws fresh points on every replay: the kernel reads the counter from device memory
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        K.sample_heat(*bufs, np.pi, 3.0, np.pi, 99, step, 0)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g, stream=s):
        K.sample_heat(*bufs, np.pi, 3.0, np.pi, 99, step, 0)
        step.add_(1)
    for it in range(3):
        g.replay()
        torch.cuda.synchronize()
        assert np.array_equal(bufs[0].cpu().numpy(), PH.heat(B, np.pi, 3.0, np.pi, 99, 17 + it)[0])
    with pytest.raises(K.DgmkError):
        K.sample_uniform(torch.zeros(4), 0.0, 1.0, 0)   # host tensor: no CPU path


def _dgrad_probe(lib, engine, A, Bt3, M, K, ld):
    import ctypes as C
    lib.dgmk_gemm_tc_probe.restype = C.c_int
    lib.dgmk_gemm_tc_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_void_p]
    lib.dgmk_set_gemm_engine(engine)
    try:
        out = torch.full((M, ld), 7.0, device="cuda")
        rc = lib.dgmk_gemm_tc_probe(A.data_ptr(), Bt3.data_ptr(), out.data_ptr(), M, 128, K, ld, None)
        torch.cuda.synchronize()
    finally:
        lib.dgmk_set_gemm_engine(1)
    assert rc == 0, lib.dgmk_last_error()
    assert bool((out[:, 128:] == 7.0).all()), "wrote outside the result block"
    return out[:, :128]


@pytest.mark.parametrize("M,K", [(1, 384), (257, 384), (40000, 384), (5000, 256), (333, 160)])
def test_dgrad_resident_kernel_vs_fp64(M, K):
    """The weight-resident K = 3H data-gradient kernel (csrc/dgmk_dgrad_res.cuh; the reverse of dgm_net.py:53-68's
    [Z | G | R] maps) against an FP64 product and the round-1 streaming tile: ragged row counts, several tiles per CTA."""
    from differential_equations_dnn_b200 import _cabi
    lib = _cabi.load()
    ld = 512
    g = torch.Generator(device="cuda").manual_seed(M)
    A = torch.randn(M, ld, device="cuda", generator=g)
    w = (torch.rand(128, K, device="cuda", generator=g) - 0.5) * 0.3
    hi = ((w.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    Bt3 = torch.cat([w.reshape(-1), hi.reshape(-1), (w - hi).reshape(-1)]).contiguous()
    ref = A[:, :K].double() @ w.double().t()
    for engine in (1, 3):
        out = _dgrad_probe(lib, engine, A, Bt3, M, K, ld)
        err = float((out.double() - ref).norm() / ref.norm())
        assert err < 5e-7, (engine, err)


def test_dgrad_resident_kernel_ring_hand_over_is_exact():
    """Every product is exact here (one-hot weights; A[r, k] encodes tile, chunk and position in the chunk), so a result
    names the tile / chunk / column it was read from: guards the hand-over of the raw ring (a stage released while its
    last shared-memory loads were still queued showed up as values of chunk g + 2, only in the MMA-bound steady state,
    i.e. from the second tile of a CTA on)."""
    from differential_equations_dnn_b200 import _cabi
    lib = _cabi.load()
    K, ld = 384, 512
    M = 74 * 128 * 4
    r = torch.arange(M, device="cuda")
    kk = torch.arange(K, device="cuda")
    w = torch.zeros(128, K, device="cuda")
    for n in range(128):
        w[n, 32 * (n % 12) + (3 * (n // 12) + n) % 32] = 1.0
    Bt3 = torch.cat([w.reshape(-1), w.reshape(-1), torch.zeros_like(w).reshape(-1)]).contiguous()
    for first in (r // 128, r % 128):
        A = torch.zeros(M, ld, device="cuda")
        A[:, :K] = first.float()[:, None] + 512.0 * (kk // 32).float()[None, :] + (kk % 32).float()[None, :] / 32
        ref = (A[:, :K].double() @ w.double().t()).float()
        for rep in range(3):
            out = _dgrad_probe(lib, 1, A, Bt3, M, K, ld)
            assert int((out != ref).sum()) == 0
