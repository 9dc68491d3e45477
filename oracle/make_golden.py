"""Generate tests/golden/*.npz by EXECUTING the unmodified reference.

Run in the build container only (needs /root/reference, which does not exist
on the GPU box):   python oracle/make_golden.py

Every fixture holds the flat parameter vector (reference named_parameters()
order), the inputs, and the loss / flat gradient that the reference's own
`dgm_loss_func` + `loss.backward()` produced in FP32 on CPU, plus the same in
FP64 (`*_f64`, the arbiter tier of SURVEY 8(c)).  Gradients of parameters the
reference leaves at `grad is None` are stored as 0 with `live`=0.

TEST INFRASTRUCTURE ONLY (see oracle/ref_port.py header).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_port as rp  # noqa: E402


def import_reference():
    """matplotlib is not installed; the reference only needs it importable."""
    mpl = types.ModuleType("matplotlib")
    pylab = types.ModuleType("matplotlib.pylab")
    pylab.rcParams = {}
    style = types.ModuleType("matplotlib.style")
    style.use = lambda *a, **k: None
    mpl.pylab, mpl.style = pylab, style
    sys.modules.update({"matplotlib": mpl, "matplotlib.pylab": pylab,
                        "matplotlib.style": style})
    sys.path.insert(0, REF)
    import neural_networks, dgm_net, heat, simple_ode, fitzhugh_nagumo, fredholm  # noqa
    return dict(nn=neural_networks, dgm=dgm_net, heat=heat, ode=simple_ode,
                fhn=fitzhugh_nagumo, fred=fredholm)


def build(ref, spec: rp.NetSpec, seed):
    torch.manual_seed(seed)
    inv = {v: k for k, v in rp.ACT_NAMES.items()}
    if spec.kind == rp.KIND_MLP:
        net = ref["nn"].MLP(input_dim=spec.d, output_dim=spec.o, hidden_size=spec.H,
                            num_layers=spec.L, activation=inv[spec.act])
    elif spec.kind == rp.KIND_DGM_LINEAR:
        net = ref["dgm"].DGM(input_dim=spec.d, output_dim=spec.o, hidden_size=spec.H,
                             num_layers=spec.L)
    else:
        net = ref["nn"].DGM(input_dim=spec.d, output_dim=spec.o, hidden_size=spec.H,
                            num_layers=spec.L, func=inv[spec.act])
    named = list(net.named_parameters())
    ent = spec.entries()
    assert [n for n, _ in named] == [e[0] for e in ent], "layout order mismatch"
    assert [tuple(p.shape) for _, p in named] == [e[1] for e in ent]
    # DGM_RAW biases start at exactly 0 (ReLU ties); perturb half the fixtures'
    # biases? No: keep the reference init, ties are part of the contract.
    return net


def flat(net):
    return torch.cat([p.detach().reshape(-1) for p in net.parameters()])


def flat_grad(net):
    gs, live = [], []
    for p in net.parameters():
        if p.grad is None:
            gs.append(torch.zeros(p.numel(), dtype=p.dtype))
            live.append(torch.zeros(p.numel(), dtype=torch.bool))
        else:
            gs.append(p.grad.detach().reshape(-1))
            live.append(torch.ones(p.numel(), dtype=torch.bool))
    return torch.cat(gs), torch.cat(live)


def spec_arr(spec):
    return np.array([spec.kind, spec.d, spec.o, spec.H, spec.L, spec.act], dtype=np.int64)


def run_both(net, fn, keep64=False):
    """fn(net, cast) -> loss ; run in fp32 then fp64."""
    out = {}
    for tag, dt in (("", torch.float32), ("_f64", torch.float64)):
        n = net.double() if dt == torch.float64 else net.float()
        n.zero_grad(set_to_none=True)
        loss = fn(n, dt)
        loss.backward()
        g, live = flat_grad(n)
        out["loss" + tag] = loss.detach().numpy()
        if dt == torch.float32 or g.numel() <= 100_000 or keep64:  # keep big fixtures small
            out["grad" + tag] = g.numpy()
        out["live"] = live.numpy()
    net.float()
    return out


def heat_inputs(B, gen):
    x = torch.pi * torch.rand([B, 1], generator=gen)
    t = 3.0 * torch.rand([B, 1], generator=gen)
    X = torch.cat([x, t], 1)
    X0 = torch.cat([x, torch.zeros(B, 1)], 1)
    XBD1 = torch.cat([torch.zeros(B, 1), t], 1)
    XBD2 = torch.cat([torch.ones(B, 1) * torch.pi, t], 1)
    return X, X0, XBD1, XBD2, torch.zeros(B, 1), torch.zeros(B, 1)


ONLY = None  # `python oracle/make_golden.py NAME [NAME...]` regenerates just those fixtures


def save(name, **kw):
    if ONLY and name not in ONLY:
        return
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **kw)
    print(f"{name}: {os.path.getsize(path)/1024:.0f} KB")


def main():
    ref = import_reference()
    S = rp.NetSpec
    MLP, DGL, DGR = rp.KIND_MLP, rp.KIND_DGM_LINEAR, rp.KIND_DGM_RAW

    # ---- heat (heat.py:50-95, inputs as the driver builds them :117-134) ----
    heat_cases = {
        "heat_dgm_h32l1": (S(DGL, 2, 1, 32, 1, rp.ACT_TANH), 64),
        "heat_dgm_h128l3": (S(DGL, 2, 1, 128, 3, rp.ACT_TANH), 24),
        # several 64-row tiles of the fused tcgen05 kernels plus a ragged tail (203 = 3 * 64 + 11 rows:
        # 812 jet rows / 609 companion rows)
        "heat_dgm_h128l3_b203": (S(DGL, 2, 1, 128, 3, rp.ACT_TANH), 203),
        "heat_dgm_h50l3": (S(DGL, 2, 1, 50, 3, rp.ACT_TANH), 37),
        "heat_mlp_tanh_h128l3": (S(MLP, 2, 1, 128, 3, rp.ACT_TANH), 64),
        "heat_mlp_relu_h128l3": (S(MLP, 2, 1, 128, 3, rp.ACT_RELU), 64),
        "heat_mlp_sigmoid_h50l1": (S(MLP, 2, 1, 50, 1, rp.ACT_SIGMOID), 33),
        "heat_mlp_leaky_h32l2": (S(MLP, 2, 1, 32, 2, rp.ACT_LEAKY), 64),
        "heat_dgmraw_h32l2": (S(DGR, 2, 1, 32, 2, rp.ACT_RELU), 64),
        # func="tanh": tanh input layer, ReLU gates (neural_networks.py:145-156)
        "heat_dgmraw_tanh_h32l1": (S(DGR, 2, 1, 32, 1, rp.ACT_TANH), 40),
    }
    for name, (spec, B) in heat_cases.items():
        if ONLY and name not in ONLY:
            continue
        net = build(ref, spec, 1234)
        gen = torch.Generator().manual_seed(1)
        inp = heat_inputs(B, gen)

        def fn(n, dt, inp=inp):
            a = [z.to(dt) for z in inp]
            a[0] = a[0].clone().requires_grad_(True)
            return ref["heat"].dgm_loss_func(n, *a)
        r = run_both(net, fn, keep64=name.endswith("_b203"))
        save(name, spec=spec_arr(spec), theta=flat(net).numpy(),
             X=inp[0].numpy(), X0=inp[1].numpy(), XBD1=inp[2].numpy(),
             XBD2=inp[3].numpy(), x_bd1=inp[4].numpy(), x_bd2=inp[5].numpy(), **r)

    # ---- simple ODE (simple_ode.py:41-63, driver :87-101) ----
    ode_cases = {
        "ode_mlp_relu_h32l1": (S(MLP, 1, 1, 32, 1, rp.ACT_RELU), 64),
        "ode_mlp_tanh_h32l1": (S(MLP, 1, 1, 32, 1, rp.ACT_TANH), 64),
        "ode_dgm_h16l2": (S(DGL, 1, 1, 16, 2, rp.ACT_TANH), 50),
    }
    for name, (spec, B) in ode_cases.items():
        net = build(ref, spec, 1234)
        gen = torch.Generator().manual_seed(1)
        t = 1.01 * torch.rand([B, 1], generator=gen)
        t0 = torch.zeros(B, 1)
        y_ic = torch.ones(B, 1) * 2.0

        def fn(n, dt):
            tt = t.to(dt).clone().requires_grad_(True)
            return ref["ode"].dgm_loss_func(n(tt), n(t0.to(dt)), tt, y_ic.to(dt))
        r = run_both(net, fn)
        save(name, spec=spec_arr(spec), theta=flat(net).numpy(), t=t.numpy(),
             t0=t0.numpy(), y_ic=y_ic.numpy(), **r)

    # ---- FitzHugh-Nagumo (fitzhugh_nagumo.py:53-97) ----
    fhn_cases = {
        "fhn_mlp_tanh_h128l3": (S(MLP, 1, 2, 128, 3, rp.ACT_TANH), 64),
        "fhn_dgm_h64l2": (S(DGL, 1, 2, 64, 2, rp.ACT_TANH), 48),
        "fhn_dgm_h128l4": (S(DGL, 1, 2, 128, 4, rp.ACT_TANH), 16),
    }
    for name, (spec, B) in fhn_cases.items():
        net = build(ref, spec, 1234)
        gen = torch.Generator().manual_seed(1)
        t = 30.01 * torch.rand([B, 1], generator=gen)
        t0 = torch.zeros(B, 1)
        y_ic = torch.zeros(B, 2)

        def fn(n, dt):
            tt = t.to(dt).clone().requires_grad_(True)
            return ref["fhn"].dgm_loss_func(n(tt), n(t0.to(dt)), tt, y_ic.to(dt))
        r = run_both(net, fn)
        save(name, spec=spec_arr(spec), theta=flat(net).numpy(), t=t.numpy(),
             t0=t0.numpy(), y_ic=y_ic.numpy(), **r)

    # ---- Fredholm (fredholm.py:47-74): nodes replayed from the RNG stream ----
    fred_cases = {
        "fredholm_dgmraw_h32l1_k50": (S(DGR, 1, 1, 32, 1, rp.ACT_RELU), 32, 50),
        "fredholm_dgmraw_h32l1_k7": (S(DGR, 1, 1, 32, 1, rp.ACT_RELU), 19, 7),
        "fredholm_mlp_tanh_h16l1_k5": (S(MLP, 1, 1, 16, 1, rp.ACT_TANH), 8, 5),
    }
    for name, (spec, B, k) in fred_cases.items():
        net = build(ref, spec, 1234)
        gen = torch.Generator().manual_seed(1)
        x = (np.pi / 2.0) * torch.rand([B, 1], generator=gen)
        torch.manual_seed(2)
        T = torch.stack([np.pi / 2.0 * torch.rand_like(x) for _ in range(k)])

        def fn(n, dt):
            torch.manual_seed(2)  # the k rand_like draws inside the reference
            if dt == torch.float64:
                # rand_like in fp64 draws different bits; feed fp32 nodes by
                # monkeypatching rand_like for this call only.
                it = iter(T.double() / (np.pi / 2.0))
                orig = torch.rand_like
                torch.rand_like = lambda z: next(it)
                try:
                    return ref["fred"].dgm_loss_func(n, x.to(dt), k)
                finally:
                    torch.rand_like = orig
            return ref["fred"].dgm_loss_func(n, x, k)
        r = run_both(net, fn)
        save(name, spec=spec_arr(spec), theta=flat(net).numpy(), x=x.numpy(),
             T=T.numpy(), **r)

    # ---- S1 seam: value / Jacobian / Hessian of the reference modules ----
    jet_cases = {
        "jets_dgm_d2_h32l1": (S(DGL, 2, 1, 32, 1, rp.ACT_TANH), 16),
        "jets_mlp_tanh_d2_o2_h16l1": (S(MLP, 2, 2, 16, 1, rp.ACT_TANH), 9),
        "jets_mlp_sigmoid_d1_h16l2": (S(MLP, 1, 1, 16, 2, rp.ACT_SIGMOID), 11),
    }
    for name, (spec, B) in jet_cases.items():
        net = build(ref, spec, 1234).double()
        gen = torch.Generator().manual_seed(3)
        X = (2.0 * torch.rand([B, spec.d], generator=gen)).double().requires_grad_(True)
        y = net(X)
        J = torch.zeros(B, spec.o, spec.d, dtype=torch.float64)
        Hs = torch.zeros(B, spec.o, spec.d, spec.d, dtype=torch.float64)
        for m in range(spec.o):
            g = torch.autograd.grad(y[:, m].sum(), X, create_graph=True)[0]
            J[:, m] = g.detach()
            for i in range(spec.d):
                Hs[:, m, i] = torch.autograd.grad(g[:, i].sum(), X, retain_graph=True)[0]
        save(name, spec=spec_arr(spec), theta=flat(net).numpy(),
             X=X.detach().numpy(), y=y.detach().numpy(), J=J.numpy(), Hs=Hs.numpy())

    # ---- Adam (heat.py:115 defaults): 5 steps of torch.optim.Adam ----
    torch.manual_seed(7)
    p = torch.nn.Parameter(torch.randn(257))
    dead = torch.nn.Parameter(torch.randn(5))  # grad None -> skipped
    opt = torch.optim.Adam([p, dead], lr=1e-3)
    grads, thetas = [], [torch.cat([p.detach(), dead.detach()]).clone().numpy()]
    for s in range(5):
        opt.zero_grad()
        g = torch.randn(257) * (10.0 ** (s - 2))
        p.grad = g.clone()
        opt.step()
        grads.append(torch.cat([g, torch.zeros(5)]).numpy())
        thetas.append(torch.cat([p.detach(), dead.detach()]).clone().numpy())
    st = opt.state[p]
    save("adam_5steps", thetas=np.stack(thetas), grads=np.stack(grads), lr=1e-3,
         live=np.array([1] * 257 + [0] * 5, dtype=np.uint8),
         m=np.concatenate([st["exp_avg"].numpy(), np.zeros(5, np.float32)]),
         v=np.concatenate([st["exp_avg_sq"].numpy(), np.zeros(5, np.float32)]))

    # ---- short end-to-end trajectory of the reference driver (simple_ode) ----
    torch.manual_seed(0)
    net = ref["nn"].MLP(input_dim=1, output_dim=1, hidden_size=32)
    theta0 = flat(net).numpy().copy()
    torch.manual_seed(5)
    _, losses = ref["ode"].minimize_loss_dgm(net, y_ic=2.0, iterations=30,
                                             batch_size=64, lrate=1e-4)
    torch.manual_seed(5)
    ts = np.stack([(1.01 * torch.rand([64, 1])).numpy() for _ in range(30)])
    save("ode_driver_30its", theta0=theta0, theta_end=flat(net).numpy(),
         losses=np.array(losses), ts=ts,
         spec=spec_arr(S(MLP, 1, 1, 32, 1, rp.ACT_RELU)))


if __name__ == "__main__":
    ONLY = set(sys.argv[1:]) or None
    main()
