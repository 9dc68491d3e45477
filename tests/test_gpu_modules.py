"""-m gpu: the drop-in Python layer (modules, autograd seams, losses, optimizer, drivers)
on top of the CUDA library, against the executed-reference golden vectors."""
import numpy as np
import pytest
import torch

from conftest import golden, golden_names, grad_groups, rel

pytestmark = pytest.mark.gpu
TOL = 1e-5


def build_net(g, seed=1234):
    from differential_equations_dnn_b200 import neural_networks as nn_, dgm_net
    kind, d, o, H, L, act = (int(v) for v in g["spec"])
    torch.manual_seed(seed)
    if kind == 0:
        net = nn_.MLP(d, o, H, L, activation={0: "relu", 1: "sigmoid", 2: "tanh", 3: "leaky_relu"}[act])
    elif kind == 1:
        net = dgm_net.DGM(d, o, H, L)
    else:
        net = nn_.DGM(d, o, H, L, func="relu" if act == 0 else "tanh")
    assert np.array_equal(net.flat_theta().numpy(), g["theta"]), "same seed must give the reference's weights"
    return net.cuda()


def reference_modules():
    """The unmodified reference modules (oracle/_ref, placed there by oracle/vendor_ref.py; they travel to the GPU
    box as a build output)."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref is empty (run __graft_entry__.build() where /root/reference exists)")
    return ref_loader.load()


def cu(g, *keys):
    return [torch.from_numpy(g[k]).cuda() for k in keys]


def check_grads(net, g, tol=TOL):
    off = 0
    layout, grads = [], []
    for name, p in net.named_parameters():
        n = p.numel()
        live = bool(g["live"][off])
        if not live:
            assert p.grad is None, name            # reference leaves grad None (dgm1.*)
        layout.append((off, n, live))
        grads.append(np.zeros(n, np.float32) if p.grad is None else p.grad.reshape(-1).cpu().numpy())
        off += n
    mine = np.concatenate(grads)
    names = [nm for nm, _ in net.named_parameters()]
    for goff, n, live in grad_groups(layout):   # norm-wise per tensor (output bias with its layer: conftest.grad_groups)
        if not live:
            continue
        ref = g["grad"][goff:goff + n]
        name = names[[o for o, _, _ in layout].index(goff)]
        if np.linalg.norm(ref) == 0:
            assert np.abs(mine[goff:goff + n]).max() < 1e-7, name
        else:
            assert rel(mine[goff:goff + n], ref) < tol, name


HEAT_KEYS = ("X", "X0", "XBD1", "XBD2", "x_bd1", "x_bd2")


@pytest.mark.parametrize("name", [n for n in golden_names("heat_")])
def test_heat_fast_path(name):
    from differential_equations_dnn_b200 import heat
    g = golden(name)
    net = build_net(g)
    loss = heat.dgm_loss_func(net, *cu(g, *HEAT_KEYS))
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= TOL * abs(float(g["loss"]))
    check_grads(net, g)


@pytest.mark.parametrize("name", ["heat_dgm_h32l1", "heat_mlp_tanh_h128l3", "heat_mlp_relu_h128l3",
                                  "heat_mlp_sigmoid_h50l1", "heat_dgmraw_h32l2"])
def test_heat_reference_code_on_our_module(name):
    """Seam S1: the UNMODIFIED reference loss (oracle/_ref/heat.py::dgm_loss_func: nested autograd.grad on
    net(x), heat.py:50-95) runs on our module and gives the reference's numbers."""
    g = golden(name)
    net = build_net(g)
    a = cu(g, *HEAT_KEYS)
    a[0].requires_grad_(True)
    loss = reference_modules().heat.dgm_loss_func(net, *a)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= TOL * abs(float(g["loss"]))
    check_grads(net, g)


@pytest.mark.parametrize("name", golden_names("ode_mlp") + golden_names("ode_dgm") + golden_names("fhn_"))
@pytest.mark.parametrize("mode", ["deferred", "eager_traced", "generic"])
def test_ode_fhn_three_paths(name, mode):
    from differential_equations_dnn_b200 import simple_ode, fitzhugh_nagumo
    from differential_equations_dnn_b200._flat import deferred_forward
    g = golden(name)
    mod = simple_ode if name.startswith("ode") else fitzhugh_nagumo
    net = build_net(g)
    t, t0, y_ic = cu(g, "t", "t0", "y_ic")
    if mode == "deferred":
        with deferred_forward(net):
            y, y0 = net(t), net(t0)
        loss = mod.dgm_loss_func(y, y0, t, y_ic)
    elif mode == "eager_traced":
        t.requires_grad_(True)
        y, y0 = net(t), net(t0)
        loss = mod.dgm_loss_func(y, y0, t, y_ic)
    else:  # seam S1: the UNMODIFIED reference loss (nested autograd.grad) on our module's outputs
        t.requires_grad_(True)
        ref = reference_modules()
        refmod = ref.simple_ode if name.startswith("ode") else ref.fitzhugh_nagumo
        loss = refmod.dgm_loss_func(net(t), net(t0), t, y_ic)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= TOL * abs(float(g["loss"]))
    check_grads(net, g)


def test_operand_shapes_are_validated_and_broadcast():
    """A y_ic the reference would broadcast (Python number, 0-dim, [1,o]) gives the full-tensor result; operands
    with too few rows raise instead of being read out of bounds."""
    from differential_equations_dnn_b200 import simple_ode, heat, kernels
    from differential_equations_dnn_b200._cabi import DgmkError
    from differential_equations_dnn_b200._flat import deferred_forward
    g = golden("ode_mlp_tanh_h32l1")
    net = build_net(g)
    t, t0, y_ic = cu(g, "t", "t0", "y_ic")
    outs = []
    for ic in (y_ic, 2.0, torch.tensor(2.0, device="cuda"), torch.full((1, 1), 2.0, device="cuda")):
        net.zero_grad()
        with deferred_forward(net):
            y, y0 = net(t), net(t0)
        loss = simple_ode.dgm_loss_func(y, y0, t, ic)
        loss.backward()
        outs.append((loss.item(), net.flat_theta().grad if False else torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone()))
    for l, gr in outs[1:]:
        assert l == outs[0][0] and torch.equal(gr, outs[0][1])
    with pytest.raises(DgmkError):
        with deferred_forward(net):
            y, y0 = net(t), net(t0[:5])
        simple_ode.dgm_loss_func(y, y0, t, y_ic)
    h = golden("heat_dgm_h32l1")
    hnet = build_net(h)
    a = cu(h, *HEAT_KEYS)
    with pytest.raises(DgmkError):
        heat.dgm_loss_func(hnet, a[0], a[1][:7], *a[2:])
    with pytest.raises(DgmkError):
        kernels.heat_step(hnet.desc, hnet.flat_theta(), a[0], a[1], a[2], a[3], a[4][:3], a[5])
    l1 = heat.dgm_loss_func(hnet, *a[:4], 0.0, 0.0)       # scalar boundary targets
    assert abs(l1.item() - float(h["loss"])) <= TOL * abs(float(h["loss"]))


def test_no_autograd_fallback_in_the_product_losses():
    """Foreign networks and broken traces raise: the package's losses have no torch-autograd path."""
    from differential_equations_dnn_b200 import heat, simple_ode, fitzhugh_nagumo, fredholm
    from differential_equations_dnn_b200._cabi import DgmkError
    g = golden("ode_mlp_tanh_h32l1")
    net = build_net(g)
    t, t0, y_ic = cu(g, "t", "t0", "y_ic")
    t.requires_grad_(True)
    for mod in (simple_ode, fitzhugh_nagumo):
        with pytest.raises(DgmkError):
            mod.dgm_loss_func(net(t) * 1.0, net(t0) * 1.0, t, y_ic)
    foreign = torch.nn.Linear(2, 1).cuda()
    x = torch.rand(8, 2, device="cuda")
    with pytest.raises(DgmkError):
        heat.dgm_loss_func(foreign, x, x, x, x, x[:, :1], x[:, :1])
    with pytest.raises(DgmkError):
        fredholm.dgm_loss_func(torch.nn.Linear(1, 1).cuda(), x[:, :1], k=3)


@pytest.mark.parametrize("name", golden_names("fredholm_"))
def test_fredholm(name):
    from differential_equations_dnn_b200 import fredholm
    g = golden(name)
    net = build_net(g)
    x, T = cu(g, "x", "T")
    loss = fredholm.dgm_loss_func(net, x, k=T.shape[0], nodes=T)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= TOL * abs(float(g["loss"]))
    check_grads(net, g)
    # default path draws its own nodes: just has to run and be finite
    l2 = fredholm.dgm_loss_func(net, x, k=7)
    assert np.isfinite(l2.item())


def test_reference_driver_trajectory():
    """30 iterations of simple_ode.minimize_loss_dgm replayed with the reference's own
    sample sequence: losses and final weights track the reference (Adam included)."""
    from differential_equations_dnn_b200 import neural_networks as nn_, simple_ode
    from differential_equations_dnn_b200._flat import deferred_forward
    from differential_equations_dnn_b200.optim import FusedAdam
    g = golden("ode_driver_30its")
    torch.manual_seed(0)
    net = nn_.MLP(input_dim=1, output_dim=1, hidden_size=32)
    assert np.array_equal(net.flat_theta().numpy(), g["theta0"])
    net = net.cuda()
    opt = FusedAdam(net.parameters(), lr=1e-4)
    y_ic = torch.ones(64, 1, device="cuda") * 2.0
    t0 = torch.zeros(64, 1, device="cuda")
    losses = []
    for i in range(30):
        t = torch.from_numpy(g["ts"][i]).cuda()
        opt.zero_grad()
        with deferred_forward(net):
            y, y0 = net(t), net(t0)
        loss = simple_ode.dgm_loss_func(y, y0, t, y_ic)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert rel(np.array(losses), g["losses"]) < 1e-5
    assert rel(net.flat_theta().cpu().numpy(), g["theta_end"]) < 1e-6


def test_module_contract():
    from differential_equations_dnn_b200 import dgm_net, neural_networks as nn_
    torch.manual_seed(3)
    net = dgm_net.DGM(2, 1, 32, 2)
    keys = list(net.state_dict().keys())
    assert keys[:2] == ["S_in.weight", "S_in.bias"] and "layers.1.H_uh.weight" in keys and keys[-1] == "S_out.bias"
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.cuda()
    th = net.flat_theta()
    assert th.is_cuda and all(p.data_ptr() >= th.data_ptr() for p in net.parameters())
    x = torch.rand(5, 2, device="cuda")
    with torch.no_grad():
        y = net(x)
    assert y.shape == (5, 1)
    net2 = dgm_net.DGM(2, 1, 32, 2).cuda()
    net2.load_state_dict(sd)
    with torch.no_grad():
        assert torch.equal(net2(x), y)
    # in-place edits through a parameter are seen by the kernels (views of one buffer)
    with torch.no_grad():
        net2.S_out.bias.add_(1.0)
        assert torch.allclose(net2(x), y + 1.0, atol=1e-6)
    # 1-D input like the reference's gridEvaluation (simple_ode.py:128-131)
    m = nn_.MLP(input_dim=1, output_dim=1, hidden_size=32).cuda()
    with torch.no_grad():
        assert m(torch.ones(1, device="cuda") * 0.3).shape == (1,)
    with pytest.raises(Exception):
        dgm_net.DGM(2, 1, 8, 1)(torch.zeros(2, 2))  # CPU tensors: no fallback


def test_eval_matches_oracle():
    from oracle import ref_port as rp
    g = golden("heat_dgm_h50l3")
    net = build_net(g)
    X = torch.from_numpy(g["X"])
    ref = rp.net_forward(rp.NetSpec(*[int(v) for v in g["spec"]]), torch.from_numpy(g["theta"]), X)
    with torch.no_grad():
        assert rel(net(X.cuda()).cpu().numpy(), ref.numpy()) < TOL


def test_simple_ode_end_to_end():
    """BASELINE config 1 end to end: 5000 its x 64 (simple_ode.py defaults), MAE vs 2exp(-t)
    on 25 nodes.  Reference (CPU, seed 0): MAE 0.00253; criterion: within 1e-3 of it."""
    from differential_equations_dnn_b200 import neural_networks as nn_, simple_ode
    torch.manual_seed(0)
    net = nn_.MLP(input_dim=1, output_dim=1, hidden_size=32).cuda()
    net, losses = simple_ode.minimize_loss_dgm(net, y_ic=2.0, iterations=5000, batch_size=64, lrate=1e-4)
    assert len(losses) == 5000 and losses[-1] < losses[0]
    sol = simple_ode.gridEvaluation(net, nodes=25)
    mae = np.abs(sol - simple_ode.exact_solution(np.linspace(0, 1.0, 25))).mean()
    assert mae < 0.00253 + 1e-3, mae


def test_heat_end_to_end_dgm():
    """heat + dgm_net.DGM(2,1,32,1), 15000 its x 64, lr 1e-4 (heat.py driver), 40x40 grid.
    Reference seeds 0/1/2: MAE 2.0e-4 / 2.2e-4 / 4.1e-4; criterion 1e-3."""
    from differential_equations_dnn_b200 import dgm_net, heat
    torch.manual_seed(0)
    net = dgm_net.DGM(input_dim=2, output_dim=1, hidden_size=32, num_layers=1).cuda()
    net, losses = heat.minimize_loss_dgm(net, iterations=15000, batch_size=64, lrate=1e-4)
    sol = heat.gridEvaluation(net, nodes=40)
    err = sol - heat.exact_solution(nodes=40)
    mae, rmse = np.abs(err).mean(), np.sqrt((err ** 2).mean())
    print("heat e2e: final loss", losses[-1], "MAE", mae, "RMSE", rmse)
    assert mae < 2.0e-4 + 1e-3 and rmse < 2.5e-4 + 1e-3


def _heat_grid_error(net):
    from differential_equations_dnn_b200 import heat
    err = heat.gridEvaluation(net, nodes=40) - heat.exact_solution(nodes=40)   # heat.py:152-172, :36-47
    return float(np.abs(err).mean()), float(np.sqrt((err ** 2).mean()))


def test_heat_end_to_end_headline_net_three_seeds():
    """north_star: final-solution L2 error within 1e-3 of the reference after the same iteration count.  The headline
    network, heat + dgm_net.DGM(2,1,128,3), at the reference driver's 15000 its x 64 rows, lr 1e-4 (heat.py:178-190),
    seeds 0 / 1 / 2 (SURVEY 8c asks for >= 3 seeds there: the reference lands at RMSE 1.62e-3, MAE 1.33e-3 with seed 0,
    BASELINE.md); the CUDA-graph driver makes a run ~12 s.  Criterion on the MEAN over the seeds, and no seed diverges."""
    from differential_equations_dnn_b200 import dgm_net, heat
    maes, rmses = [], []
    for seed in (0, 1, 2):
        torch.manual_seed(seed)
        net = dgm_net.DGM(input_dim=2, output_dim=1, hidden_size=128, num_layers=3).cuda()
        net, losses = heat.minimize_loss_dgm(net, iterations=15000, batch_size=64, lrate=1e-4, cuda_graph=True)
        assert len(losses) == 15000 and np.all(np.isfinite(losses)) and np.mean(losses[-100:]) < 1e-3
        mae, rmse = _heat_grid_error(net)
        print(f"heat DGM(2,1,128,3) seed {seed}: final loss {losses[-1]:.2e} MAE {mae:.2e} RMSE {rmse:.2e}")
        maes.append(mae); rmses.append(rmse)
    assert np.mean(rmses) < 1.62e-3 + 1e-3 and np.mean(maes) < 1.33e-3 + 1e-3, (maes, rmses)
    assert max(rmses) < 1e-2, rmses


def test_heat_end_to_end_tanh_mlp():
    """heat + MLP(2,1,128,3, tanh) (the paper's network), 15000 its x 64: reference MAE 3.1e-4 / RMSE 3.8e-4."""
    from differential_equations_dnn_b200 import neural_networks as nn_, heat
    torch.manual_seed(0)
    net = nn_.MLP(input_dim=2, output_dim=1, hidden_size=128, num_layers=3, activation="tanh").cuda()
    net, losses = heat.minimize_loss_dgm(net, iterations=15000, batch_size=64, lrate=1e-4, cuda_graph=True)
    mae, rmse = _heat_grid_error(net)
    print(f"heat MLP tanh: final loss {losses[-1]:.2e} MAE {mae:.2e} RMSE {rmse:.2e}")
    assert mae < 3.1e-4 + 1e-3 and rmse < 3.8e-4 + 1e-3, (mae, rmse)


def test_fredholm_end_to_end():
    """fredholm.py as shipped: neural_networks.DGM(1,1,32), 3000 its x 32 rows, k = 50, lr 1e-4 (fredholm.py:143-181);
    MAE vs 2 sin x on 50 nodes of [0, pi/2].  The loss is a Monte-Carlo estimate (k = 50 nodes per point) and the result
    depends on the seed: the UNMODIFIED reference driver (oracle/_ref, CPU, torch.manual_seed(s)) gives, for s = 0..3,
    MAE 0.00932 / 0.01040 / 0.01316 / 0.00626 and RMSE 0.01082 / 0.01162 / 0.01543 / 0.00787 (means 0.00979 / 0.01144;
    seed 0 is the BASELINE.md row).  The GPU sampler draws a different stream, so three seeds are run here and the
    criterion -- within 1e-3 of the reference -- is applied mean against mean."""
    from differential_equations_dnn_b200 import neural_networks as nn_, fredholm
    maes, rmses = [], []
    for seed in (0, 1, 2):
        torch.manual_seed(seed)
        net = nn_.DGM(input_dim=1, output_dim=1, hidden_size=32).cuda()
        net, losses = fredholm.minimize_loss_dgm(net, iterations=3000, batch_size=32, lrate=1e-4, k=50, cuda_graph=True)
        assert len(losses) == 3000 and np.all(np.isfinite(losses))
        sol = fredholm.gridEvaluation(net, nodes=50)
        err = sol - fredholm.exact_solution(np.linspace(0, np.pi / 2.0, 50))
        maes.append(float(np.abs(err).mean())); rmses.append(float(np.sqrt((err ** 2).mean())))
        print(f"fredholm seed {seed}: final loss {losses[-1]:.2e} MAE {maes[-1]:.2e} RMSE {rmses[-1]:.2e}")
    assert np.mean(maes) < 0.00979 + 1e-3 and np.mean(rmses) < 0.01144 + 1e-3, (maes, rmses)
    assert max(rmses) < 0.01543 + 1e-3, rmses   # no seed worse than the reference's worst


def test_trial_launcher_matches_serial_objective():
    """N3: `parallel.run_trials` with the port of objectiveRay (optimize_heat_ray.py:133-157) returns, on one GPU, the
    losses of calling the objective serially; the batch-size study (batchsize_effect_heat.py:186-202) runs its trials
    through the same launcher, quirks of the shipped loop included (every curve trains with batch 64)."""
    from differential_equations_dnn_b200 import optimize_heat_ray as ohr, batchsize_effect_heat as bse, parallel
    configs = [dict(c, n_iters=150) for c in parallel.sample_search_space(3, 0)]
    assert all(1 <= c["batch_size"] < 512 and 1e-4 <= c["lrate"] <= 1e-1 for c in configs)

    def obj(cfg):
        torch.manual_seed(5)
        torch.cuda.manual_seed(5)
        return ohr.objectiveRay(cfg)
    res = parallel.run_trials(obj, configs)
    serial = [obj(c) for c in configs]
    assert [r["loss"] for r in res] == serial and all(np.isfinite(serial))
    assert parallel.best_trial(res)["loss"] == min(serial)
    torch.manual_seed(6)
    study = bse.run_study(n_iters=120, n_runs=2, n_batches=2)
    assert [r["config"]["batch_size"] for r in study] == [1, 2, 4] and all(np.isfinite(r["loss"]) for r in study)
    torch.manual_seed(6)
    fixed = bse.run_study(n_iters=120, n_runs=2, n_batches=2, fix_batch_size=True, fresh_net=True)
    assert len(fixed) == 3 and all(np.isfinite(r["loss"]) for r in fixed)


def test_cuda_graph_driver_matches_eager():
    """heat.minimize_loss_dgm(cuda_graph=True) -- one captured iteration (device sampler, fused step,
    Adam with a device-resident step counter) replayed -- follows the eager loop: same RNG stream, same
    arithmetic, so the loss trajectories agree (chaotic divergence aside: early iterations tight,
    the whole curve statistically) and the final weights are finite."""
    from differential_equations_dnn_b200 import dgm_net, heat
    its = 60
    curves = []
    for graph in (False, True):
        torch.manual_seed(1234)
        net = dgm_net.DGM(input_dim=2, output_dim=1, hidden_size=32, num_layers=1).cuda()
        torch.manual_seed(99)
        torch.cuda.manual_seed(99)
        _, loss = heat.minimize_loss_dgm(net, iterations=its, batch_size=64, lrate=1e-3, cuda_graph=graph)
        assert len(loss) == its and np.all(np.isfinite(loss))
        assert torch.isfinite(net.flat_theta()).all()
        curves.append(np.array(loss))
    eager, graphed = curves
    assert np.allclose(eager[:15], graphed[:15], rtol=1e-4), (eager[:15], graphed[:15])
    assert abs(eager[-10:].mean() - graphed[-10:].mean()) <= 0.2 * abs(eager[-10:].mean())
    assert graphed[-1] < graphed[0]


@pytest.mark.parametrize("problem", ["simple_ode", "fhn_uniform", "fredholm"])
def test_cuda_graph_driver_other_problems(problem):
    """The same replay driver behind the other three `minimize_loss_dgm` (simple_ode.py:66-112,
    fitzhugh_nagumo.py:100-156 with the uniform sampler, fredholm.py:77-117): the graphed loop follows
    the eager one (same RNG stream and arithmetic)."""
    from differential_equations_dnn_b200 import neural_networks, simple_ode, fitzhugh_nagumo, fredholm
    its = 40
    curves = []
    for graph in (False, True):
        torch.manual_seed(1234)
        if problem == "simple_ode":
            net = neural_networks.MLP(input_dim=1, output_dim=1, hidden_size=32).cuda()
        elif problem == "fhn_uniform":
            net = neural_networks.MLP(input_dim=1, output_dim=2, hidden_size=32, num_layers=2, activation="tanh").cuda()
        else:
            net = neural_networks.DGM(input_dim=1, output_dim=1, hidden_size=32).cuda()
        torch.manual_seed(7)
        torch.cuda.manual_seed(7)
        if problem == "simple_ode":
            _, loss = simple_ode.minimize_loss_dgm(net, y_ic=2.0, iterations=its, batch_size=64, lrate=1e-3, cuda_graph=graph)
        elif problem == "fhn_uniform":
            y_ic = torch.zeros([64, 2], device="cuda")
            _, loss = fitzhugh_nagumo.minimize_loss_dgm(net, y_ic, iterations=its, batch_size=64, lrate=1e-3,
                                                        sampler="uniform", cuda_graph=graph)
        else:
            _, loss = fredholm.minimize_loss_dgm(net, iterations=its, batch_size=32, lrate=1e-3, k=8, cuda_graph=graph)
        assert len(loss) == its and np.all(np.isfinite(loss))
        assert torch.isfinite(net.flat_theta()).all()
        curves.append(np.array(loss))
    eager, graphed = curves
    assert np.allclose(eager[:15], graphed[:15], rtol=1e-4), (eager[:15], graphed[:15])
    assert abs(eager[-10:].mean() - graphed[-10:].mean()) <= 0.25 * abs(eager[-10:].mean()) + 1e-6


@pytest.mark.parametrize("kind", ["dgm", "mlp"])
def test_eval_and_jets_hidden128(kind):
    """Hidden size 128 through the module seams: value-only evaluation (fused forward kernels, ragged
    row count) and the generic jet seam (channel set of 6: stays on the un-fused kernels) against the
    FP64 jet oracle."""
    from oracle import jets_np
    from differential_equations_dnn_b200 import dgm_net, neural_networks
    torch.manual_seed(5)
    net = (dgm_net.DGM(2, 1, 128, 2) if kind == "dgm" else neural_networks.MLP(2, 1, 128, 2, activation="tanh")).cuda()
    d = net.desc
    spec = np.array([d.kind, d.input_dim, d.output_dim, d.hidden_size, d.num_layers, d.activation])
    gen = torch.Generator().manual_seed(6)
    X = torch.rand(1000 + 13, 2, generator=gen) * torch.tensor([np.pi, 3.0])
    y64, J64, H64 = jets_np.jets_full(spec, net.flat_theta().double().cpu().numpy(), X.double().numpy())
    with torch.no_grad():
        y = net(X.cuda()).double().cpu().numpy()
    assert rel(y, y64) < TOL
    Xg = X.cuda().requires_grad_(True)
    u = net(Xg)
    (J,) = torch.autograd.grad(u, Xg, grad_outputs=torch.ones_like(u), create_graph=True)
    assert rel(J.detach().double().cpu().numpy(), J64[:, 0, :]) < TOL
    (Jx,) = torch.autograd.grad(J[:, 0], Xg, grad_outputs=torch.ones_like(J[:, 0]), create_graph=True)
    assert rel(Jx[:, 0].detach().double().cpu().numpy(), H64[:, 0, 0, 0]) < TOL


@pytest.mark.parametrize("problem", ["heat", "simple_ode", "fhn", "fredholm"])
def test_philox_sampler_drivers(problem):
    """sampler="philox" (SURVEY 8f N2): the collocation points of every `minimize_loss_dgm` come from ONE launch of this
    library's counter-based sampler.  A draw is a pure function of (seed, stream, iteration), so the eager loop
    (iteration passed from the host) and the CUDA-graph loop (iteration read from the device counter the replay
    advances) see the same points: the trajectories agree like the torch-sampler ones do, and training makes
    progress."""
    from differential_equations_dnn_b200 import dgm_net, neural_networks, heat, simple_ode, fitzhugh_nagumo, fredholm
    its = 60
    curves = []
    for graph in (False, True):
        torch.manual_seed(1234)
        if problem == "heat":
            net = dgm_net.DGM(input_dim=2, output_dim=1, hidden_size=32, num_layers=1).cuda()
            _, loss = heat.minimize_loss_dgm(net, iterations=its, batch_size=64, lrate=1e-3, cuda_graph=graph, sampler="philox")
        elif problem == "simple_ode":
            net = neural_networks.MLP(input_dim=1, output_dim=1, hidden_size=32).cuda()
            _, loss = simple_ode.minimize_loss_dgm(net, y_ic=2.0, iterations=its, batch_size=64, lrate=1e-3, cuda_graph=graph,
                                                   sampler="philox")
        elif problem == "fhn":
            net = neural_networks.MLP(input_dim=1, output_dim=2, hidden_size=32, num_layers=2, activation="tanh").cuda()
            _, loss = fitzhugh_nagumo.minimize_loss_dgm(net, torch.zeros([64, 2], device="cuda"), iterations=its, batch_size=64,
                                                        lrate=1e-3, sampler="philox", cuda_graph=graph)
        else:
            net = neural_networks.DGM(input_dim=1, output_dim=1, hidden_size=32).cuda()
            _, loss = fredholm.minimize_loss_dgm(net, iterations=its, batch_size=32, lrate=1e-3, k=8, cuda_graph=graph,
                                                 sampler="philox")
        assert len(loss) == its and np.all(np.isfinite(loss))
        assert torch.isfinite(net.flat_theta()).all()
        curves.append(np.array(loss))
    eager, graphed = curves
    assert np.allclose(eager[:15], graphed[:15], rtol=1e-4), (eager[:15], graphed[:15])
    assert abs(eager[-10:].mean() - graphed[-10:].mean()) <= 0.25 * abs(eager[-10:].mean()) + 1e-6
    assert graphed[-10:].mean() < graphed[:10].mean()


def test_simple_ode_end_to_end_philox_graph():
    """BASELINE config 1 end to end with the on-device sampler and the CUDA-graph driver: 5000 its x 64, MAE vs
    2 exp(-t) on 25 nodes within 1e-3 of the reference's 0.00253 (statistical parity of the sampler)."""
    from differential_equations_dnn_b200 import neural_networks as nn_, simple_ode
    torch.manual_seed(0)
    net = nn_.MLP(input_dim=1, output_dim=1, hidden_size=32).cuda()
    net, losses = simple_ode.minimize_loss_dgm(net, y_ic=2.0, iterations=5000, batch_size=64, lrate=1e-4, cuda_graph=True,
                                               sampler="philox")
    assert len(losses) == 5000 and losses[-1] < losses[0]
    sol = simple_ode.gridEvaluation(net, nodes=25)
    mae = np.abs(sol - simple_ode.exact_solution(np.linspace(0, 1.0, 25))).mean()
    assert mae < 0.00253 + 1e-3, mae
