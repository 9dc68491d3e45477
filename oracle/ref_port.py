"""CPU oracle, tier 1 -- functional torch restatement of the reference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in `differential_equations_dnn_b200/` may
import this file; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` do.

What it restates (all citations into /root/reference):
  * the three network families as pure functions of a FLAT parameter vector
      MLP              neural_networks.py:230-245  (ctor :184-228)
      dgm_net.DGM      dgm_net.py:103-119, layer :53-68
      neural_networks.DGM (raw [in,out] weights, ReLU gates)
                       neural_networks.py:106-127, :162-177
  * the four residual losses, using nested torch.autograd.grad exactly like
    the reference does (this is the algorithm whose cost the CPU baseline is
    supposed to show):
      heat      heat.py:50-95
      ode       simple_ode.py:41-63
      fhn       fitzhugh_nagumo.py:53-97
      fredholm  fredholm.py:47-74   (MC nodes passed in explicitly, in the
                                     order the k rand_like calls produce them)
  * torch.optim.Adam's default single-tensor arithmetic (heat.py:115).

Parity pin: `tests/golden/*.npz` were produced by `oracle/make_golden.py`,
which imports and executes the unmodified reference from /root/reference;
`tests/test_oracle.py` checks this port against every one of them.

The flat layout is the order of `named_parameters()` of the reference
modules (checked against the real modules in make_golden.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

KIND_MLP, KIND_DGM_LINEAR, KIND_DGM_RAW = 0, 1, 2
ACT_RELU, ACT_SIGMOID, ACT_TANH, ACT_LEAKY = 0, 1, 2, 3
ACT_NAMES = {"relu": ACT_RELU, "sigmoid": ACT_SIGMOID, "tanh": ACT_TANH,
             "leaky_relu": ACT_LEAKY}


@dataclass(frozen=True)
class NetSpec:
    kind: int
    d: int
    o: int
    H: int
    L: int
    act: int = ACT_TANH  # MLP: any; DGM_LINEAR: tanh; DGM_RAW: relu

    def entries(self):
        """[(name, shape, live)] in reference named_parameters() order."""
        d, o, H, L = self.d, self.o, self.H, self.L
        out = []
        if self.kind == KIND_MLP:
            out += [("fc_in.weight", (H, d), True), ("fc_in.bias", (H,), True)]
            for i in range(L):
                out += [(f"layers.{i}.weight", (H, H), True),
                        (f"layers.{i}.bias", (H,), True)]
            out += [("fc_out.weight", (o, H), True), ("fc_out.bias", (o,), True)]
        elif self.kind == KIND_DGM_LINEAR:
            out += [("S_in.weight", (H, d), True), ("S_in.bias", (H,), True)]
            for i in range(L):
                for w, u in (("Z_wg", "Z_ug"), ("G_wz", "G_uz"),
                             ("R_wr", "R_ur"), ("H_wh", "H_uh")):
                    out += [(f"layers.{i}.{w}.weight", (H, H), True),
                            (f"layers.{i}.{w}.bias", (H,), True),
                            (f"layers.{i}.{u}.weight", (H, d), True)]
            out += [("S_out.weight", (o, H), True), ("S_out.bias", (o,), True)]
        elif self.kind == KIND_DGM_RAW:
            out += [("x_in.weight", (H, d), True), ("x_in.bias", (H,), True)]

            def layer(prefix, live):
                r = []
                for n in ("Uz", "Ug", "Ur", "Uh"):
                    r.append((f"{prefix}.{n}", (d, H), live))
                for n in ("Wz", "Wg", "Wr", "Wh"):
                    r.append((f"{prefix}.{n}", (H, H), live))
                for n in ("bz", "bg", "br", "bh"):
                    r.append((f"{prefix}.{n}", (1, H), live))
                return r
            out += layer("dgm1", False)  # registered but never used (:145)
            for i in range(L):
                out += layer(f"layers.{i}", True)
            out += [("x_out.weight", (o, H), True), ("x_out.bias", (o,), True)]
        else:
            raise ValueError(self.kind)
        return out

    def num_params(self):
        return sum(math.prod(s) for _, s, _ in self.entries())

    def views(self, theta):
        """dict name -> view of flat theta."""
        v, off = {}, 0
        for name, shape, _ in self.entries():
            n = math.prod(shape)
            v[name] = theta[off:off + n].view(shape)
            off += n
        assert off == theta.numel()
        return v

    def live_mask(self):
        m = []
        for _, shape, live in self.entries():
            m += [live] * math.prod(shape)
        return torch.tensor(m, dtype=torch.bool)


def _act(spec_act, x):
    if spec_act == ACT_RELU:
        return torch.relu(x)
    if spec_act == ACT_SIGMOID:
        return torch.sigmoid(x)
    if spec_act == ACT_TANH:
        return torch.tanh(x)
    return torch.nn.functional.leaky_relu(x, 0.01)


def net_forward(spec: NetSpec, theta: torch.Tensor, x: torch.Tensor):
    """u = net(x) for x [B, d]; differentiable in theta and x."""
    p = spec.views(theta)
    if spec.kind == KIND_MLP:
        h = _act(spec.act, x @ p["fc_in.weight"].T + p["fc_in.bias"])
        for i in range(spec.L):
            h = _act(spec.act, h @ p[f"layers.{i}.weight"].T + p[f"layers.{i}.bias"])
        return h @ p["fc_out.weight"].T + p["fc_out.bias"]
    if spec.kind == KIND_DGM_LINEAR:
        s = torch.tanh(x @ p["S_in.weight"].T + p["S_in.bias"])
        for i in range(spec.L):
            q = f"layers.{i}."

            def gate(w, u, inp):
                return torch.tanh(inp @ p[q + w + ".weight"].T + p[q + w + ".bias"]
                                  + x @ p[q + u + ".weight"].T)
            Z = gate("Z_wg", "Z_ug", s)
            G = gate("G_wz", "G_uz", s)
            R = gate("R_wr", "R_ur", s)
            Hh = gate("H_wh", "H_uh", s * R)
            s = (1 - G) * Hh + Z * s
        return s @ p["S_out.weight"].T + p["S_out.bias"]
    # DGM_RAW: relu layers, [in,out] matrices (neural_networks.py:115-126); the input layer follows `func`
    # (relu as shipped, tanh otherwise: neural_networks.py:153-156,172)
    s = _act(spec.act, x @ p["x_in.weight"].T + p["x_in.bias"])
    for i in range(spec.L):
        q = f"layers.{i}."
        Z = torch.relu(x @ p[q + "Uz"] + s @ p[q + "Wz"] + p[q + "bz"])
        G = torch.relu(x @ p[q + "Ug"] + s @ p[q + "Wg"] + p[q + "bg"])
        R = torch.relu(x @ p[q + "Ur"] + s @ p[q + "Wr"] + p[q + "br"])
        Hh = torch.relu(x @ p[q + "Uh"] + (s * R) @ p[q + "Wh"] + p[q + "bh"])
        s = (1 - G) * Hh + Z * s
    return s @ p["x_out.weight"].T + p["x_out.bias"]


def _grad(y, x):
    return torch.autograd.grad(y, x, grad_outputs=torch.ones_like(y),
                               create_graph=True, retain_graph=True)[0]


def heat_loss(spec, theta, X, X0, XBD1, XBD2, x_bd1, x_bd2, kappa=1.0):
    """heat.py:50-95.  X has requires_grad set by the caller or here."""
    X = X.detach().requires_grad_(True)
    y = net_forward(spec, theta, X)
    dy = _grad(y, X)
    dydt, dydx = dy[:, 1:2], dy[:, 0:1]
    dydxx = torch.autograd.grad(dydx, X, grad_outputs=torch.ones_like(y),
                                create_graph=True, retain_graph=True)[0][:, 0:1]
    L_dom = (dydt - kappa * dydxx) ** 2
    y0 = net_forward(spec, theta, X0)
    L_init = (y0 - torch.sin(X0[:, 0:1])) ** 2
    L_bd = (net_forward(spec, theta, XBD1) - x_bd1) ** 2 \
        + (net_forward(spec, theta, XBD2) - x_bd2) ** 2
    return torch.mean(L_dom + L_init + L_bd)


def ode_loss(spec, theta, t, t0, y_ic):
    """simple_ode.py:41-63 with y=net(t), y0=net(t0) (driver :98-101)."""
    t = t.detach().requires_grad_(True)
    y = net_forward(spec, theta, t)
    y0 = net_forward(spec, theta, t0)
    dydt = _grad(y, t)
    return torch.mean((dydt + y) ** 2 + (y0 - y_ic) ** 2)


FHN_I, FHN_ALPHA, FHN_BETA, FHN_TAU = 0.5, 0.7, 0.8, 2.5


def fhn_loss(spec, theta, t, t0, y_ic):
    """fitzhugh_nagumo.py:53-97 (three separate means; L0 is over 2B elems)."""
    t = t.detach().requires_grad_(True)
    y = net_forward(spec, theta, t)
    y0 = net_forward(spec, theta, t0)
    Y, W = y[:, 0:1], y[:, 1:2]
    dY = _grad(Y, t)
    dW = _grad(W, t)
    Lx = torch.mean((dY + (Y ** 3 / 3.0 + W - FHN_I - Y)) ** 2)
    Ly = torch.mean((dW + (FHN_BETA * W - FHN_ALPHA - Y) / FHN_TAU) ** 2)
    L0 = torch.mean((y0 - y_ic) ** 2)
    return Lx + Ly + L0


def fredholm_loss(spec, theta, x, T):
    """fredholm.py:47-74; T is [k, B, 1], node j = j-th rand_like draw."""
    k = T.shape[0]
    dr = math.pi / (2 * k)
    integral = 0.0
    for j in range(k):
        integral = integral + torch.sin(x) * torch.cos(T[j]) * net_forward(spec, theta, T[j])
    integral = integral * dr
    yhat = net_forward(spec, theta, x)
    return torch.mean((yhat - torch.sin(x) - integral) ** 2)


def loss_and_grad(fn, spec, theta, *args, **kw):
    """Run one reference-style step: loss + d loss / d theta (flat)."""
    th = theta.detach().clone().requires_grad_(True)
    loss = fn(spec, th, *args, **kw)
    loss.backward()
    g = th.grad if th.grad is not None else torch.zeros_like(th)
    return loss.detach(), g.detach()


def adam_step(theta, m, v, g, step, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8, live=None):
    """torch.optim.Adam defaults, single-tensor form (torch/optim/adam.py
    `_single_tensor_adam`): step is the 1-based count AFTER increment."""
    m2 = b1 * m + (1 - b1) * g
    v2 = b2 * v + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = v2.sqrt() / math.sqrt(bc2) + eps
    th2 = theta - (lr / bc1) * m2 / denom
    if live is not None:
        th2 = torch.where(live, th2, theta)
        m2 = torch.where(live, m2, m)
        v2 = torch.where(live, v2, v)
    return th2, m2, v2


def jets_full(spec, theta, X):
    """y [B,o], J [B,o,d], Hs [B,o,d,d] by nested autograd (S1 seam oracle)."""
    X = X.detach().requires_grad_(True)
    y = net_forward(spec, theta, X)
    B, o, d = X.shape[0], spec.o, spec.d
    J = torch.zeros(B, o, d, dtype=X.dtype)
    Hs = torch.zeros(B, o, d, d, dtype=X.dtype)
    for m in range(o):
        g = torch.autograd.grad(y[:, m].sum(), X, create_graph=True)[0]
        J[:, m] = g.detach()
        for i in range(d):
            if g.requires_grad:
                h = torch.autograd.grad(g[:, i].sum(), X, retain_graph=True,
                                        allow_unused=True)[0]
                if h is not None:
                    Hs[:, m, i] = h
    return y.detach(), J, Hs
