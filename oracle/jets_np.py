"""CPU oracle, tier 2 -- numpy restatement of the jet (Taylor-mode) forward pass
and its hand-written adjoint: the ALGORITHM the sm_100a kernels implement.

TEST INFRASTRUCTURE ONLY (see oracle/ref_port.py header).

The reference obtains u_t, u_x, u_xx by nested torch.autograd.grad
(heat.py:73-85, simple_ode.py:54-58, fitzhugh_nagumo.py:74-84).  Here every
activation carries channels [v, d_0..d_{nd-1}, p_0..p_{np-1}] (value, first
derivatives along chosen input coordinates, chosen second derivatives) and the
network (neural_networks.py:230-245, dgm_net.py:53-68,103-119,
neural_networks.py:106-127,162-177) is pushed forward channel-wise; the reverse
sweep is the adjoint of that program.  SURVEY 7.1 lists the rules.

Parity pin: checked against tests/golden (executed reference) in
tests/test_oracle.py, FP64 vs the reference's FP64 run.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

KIND_MLP, KIND_DGM_LINEAR, KIND_DGM_RAW = 0, 1, 2
ACT_RELU, ACT_SIGMOID, ACT_TANH, ACT_LEAKY = 0, 1, 2, 3


@dataclass(frozen=True)
class ChanSet:
    coords: tuple  # input coordinate of first-order direction k
    pairs: tuple   # (i, j) direction indices, i <= j, of second-order channel q

    @property
    def nd(self):
        return len(self.coords)

    @property
    def C(self):
        return 1 + len(self.coords) + len(self.pairs)


CS_V = ChanSet((), ())
CS_D1O1 = ChanSet((0,), ())                       # ODE / FHN: v, t
CS_HEAT = ChanSet((0, 1), ((0, 0),))              # v, x, t, xx
CS_D2O1 = ChanSet((0, 1), ())
CS_D1O2 = ChanSet((0,), ((0, 0),))
CS_D2O2 = ChanSet((0, 1), ((0, 0), (0, 1), (1, 1)))


# ---------------------------------------------------------------- activations
def act_derivs(act, y, a=None):
    """sigma', sigma'', sigma''' expressed through the OUTPUT y (a-form stash)."""
    if act == ACT_TANH:
        d1 = 1 - y * y
        return d1, -2 * y * d1, d1 * (6 * y * y - 2)
    if act == ACT_SIGMOID:
        d1 = y * (1 - y)
        return d1, d1 * (1 - 2 * y), d1 * (1 - 6 * d1)
    z = np.zeros_like(y)
    if act == ACT_RELU:
        return (y > 0).astype(y.dtype), z, z
    return np.where(y > 0, 1.0, 0.01).astype(y.dtype), z, z


def act_value(act, a):
    if act == ACT_TANH:
        return np.tanh(a)
    if act == ACT_SIGMOID:
        return 1 / (1 + np.exp(-a))
    if act == ACT_RELU:
        return np.maximum(a, 0)
    return np.where(a > 0, a, 0.01 * a)


def act_fwd(cs, act, a):
    """a: [C, ...] pre-activation jet -> y jet, plus a-form stash (y_v, a_1..)."""
    y = np.empty_like(a)
    y[0] = act_value(act, a[0])
    d1, d2, _ = act_derivs(act, y[0])
    for k in range(cs.nd):
        y[1 + k] = d1 * a[1 + k]
    for q, (i, j) in enumerate(cs.pairs):
        c = 1 + cs.nd + q
        y[c] = d1 * a[c] + d2 * a[1 + i] * a[1 + j]
    return y


def act_adj(cs, act, ybar, yv, a):
    """Cotangent of the pre-activation jet. a[0] is unused (yv carries it)."""
    d1, d2, d3 = act_derivs(act, yv)
    abar = np.zeros_like(ybar)
    abar[0] = d1 * ybar[0]
    for k in range(cs.nd):
        abar[1 + k] = d1 * ybar[1 + k]
        abar[0] += d2 * a[1 + k] * ybar[1 + k]
    for q, (i, j) in enumerate(cs.pairs):
        c = 1 + cs.nd + q
        abar[c] = d1 * ybar[c]
        abar[1 + i] += d2 * a[1 + j] * ybar[c]
        abar[1 + j] += d2 * a[1 + i] * ybar[c]
        abar[0] += (d2 * a[c] + d3 * a[1 + i] * a[1 + j]) * ybar[c]
    return abar


def prod_fwd(cs, p, q):
    r = np.empty_like(p)
    r[0] = p[0] * q[0]
    for k in range(cs.nd):
        r[1 + k] = p[1 + k] * q[0] + p[0] * q[1 + k]
    for n, (i, j) in enumerate(cs.pairs):
        c = 1 + cs.nd + n
        r[c] = p[c] * q[0] + p[1 + i] * q[1 + j] + p[1 + j] * q[1 + i] + p[0] * q[c]
    return r


def prod_adj(cs, rbar, q):
    """Cotangent w.r.t. p of r = p*q (q is the other factor's jet)."""
    pbar = np.zeros_like(rbar)
    pbar[0] = q[0] * rbar[0]
    for k in range(cs.nd):
        pbar[0] += q[1 + k] * rbar[1 + k]
        pbar[1 + k] = q[0] * rbar[1 + k]
    for n, (i, j) in enumerate(cs.pairs):
        c = 1 + cs.nd + n
        pbar[0] += q[c] * rbar[c]
        pbar[1 + i] += q[1 + j] * rbar[c]
        pbar[1 + j] += q[1 + i] * rbar[c]
        pbar[c] = q[0] * rbar[c]
    return pbar


# ---------------------------------------------------------------- layout
def entries(kind, d, o, H, L):
    out = []
    if kind == KIND_MLP:
        out += [("fc_in.weight", (H, d)), ("fc_in.bias", (H,))]
        for i in range(L):
            out += [(f"layers.{i}.weight", (H, H)), (f"layers.{i}.bias", (H,))]
        out += [("fc_out.weight", (o, H)), ("fc_out.bias", (o,))]
    elif kind == KIND_DGM_LINEAR:
        out += [("S_in.weight", (H, d)), ("S_in.bias", (H,))]
        for i in range(L):
            for w, u in (("Z_wg", "Z_ug"), ("G_wz", "G_uz"), ("R_wr", "R_ur"), ("H_wh", "H_uh")):
                out += [(f"layers.{i}.{w}.weight", (H, H)), (f"layers.{i}.{w}.bias", (H,)),
                        (f"layers.{i}.{u}.weight", (H, d))]
        out += [("S_out.weight", (o, H)), ("S_out.bias", (o,))]
    else:
        out += [("x_in.weight", (H, d)), ("x_in.bias", (H,))]
        for pre in ["dgm1"] + [f"layers.{i}" for i in range(L)]:
            out += [(f"{pre}.{n}", (d, H)) for n in ("Uz", "Ug", "Ur", "Uh")]
            out += [(f"{pre}.{n}", (H, H)) for n in ("Wz", "Wg", "Wr", "Wh")]
            out += [(f"{pre}.{n}", (1, H)) for n in ("bz", "bg", "br", "bh")]
        out += [("x_out.weight", (o, H)), ("x_out.bias", (o,))]
    return out


def views(kind, d, o, H, L, theta):
    v, off = {}, 0
    for name, shape in entries(kind, d, o, H, L):
        n = math.prod(shape)
        v[name] = theta[off:off + n].reshape(shape)
        off += n
    assert off == theta.size
    return v


class Net:
    """Canonical per-layer view: every linear map as (W [out,in], U [out,d], b [out])."""

    def __init__(self, kind, d, o, H, L, act, theta):
        self.kind, self.d, self.o, self.H, self.L, self.act = kind, d, o, H, L, act
        self.theta = theta
        p = views(kind, d, o, H, L, theta)
        self.p = p
        if kind == KIND_MLP:
            self.Win, self.bin = p["fc_in.weight"], p["fc_in.bias"]
            self.Wout, self.bout = p["fc_out.weight"], p["fc_out.bias"]
            self.hidden = [(p[f"layers.{i}.weight"], p[f"layers.{i}.bias"]) for i in range(L)]
            self.gact = act
        elif kind == KIND_DGM_LINEAR:
            self.Win, self.bin = p["S_in.weight"], p["S_in.bias"]
            self.Wout, self.bout = p["S_out.weight"], p["S_out.bias"]
            self.gates = []
            for i in range(L):
                q = f"layers.{i}."
                self.gates.append({g: (p[q + w + ".weight"], p[q + u + ".weight"], p[q + w + ".bias"])
                                   for g, w, u in (("Z", "Z_wg", "Z_ug"), ("G", "G_wz", "G_uz"),
                                                   ("R", "R_wr", "R_ur"), ("H", "H_wh", "H_uh"))})
            self.act = self.gact = ACT_TANH
        else:
            self.Win, self.bin = p["x_in.weight"], p["x_in.bias"]
            self.Wout, self.bout = p["x_out.weight"], p["x_out.bias"]
            self.gates = []
            for i in range(L):
                q = f"layers.{i}."
                self.gates.append({g: (p[q + "W" + s].T, p[q + "U" + s].T, p[q + "b" + s][0])
                                   for g, s in (("Z", "z"), ("G", "g"), ("R", "r"), ("H", "h"))})
            self.gact = ACT_RELU   # self.act (input layer) = func: relu as shipped, or tanh

    # gradient accumulation mirrors the canonical view back to the flat layout
    def grad_views(self, g):
        return Net(self.kind, self.d, self.o, self.H, self.L, self.act, g)


def input_jet(cs, X):
    """[C, B, d] jet of the identity map."""
    B, d = X.shape
    s = np.zeros((cs.C, B, d), X.dtype)
    s[0] = X
    for k, c in enumerate(cs.coords):
        s[1 + k, :, c] = 1
    return s


def lin(cs, W, U, b, s, xj):
    """a_c = s_c W^T + x_c U^T (+ b on the value channel)."""
    a = s @ W.T
    if U is not None:
        a = a + xj @ U.T
    a[0] = a[0] + b
    return a


def forward(net: Net, cs: ChanSet, X):
    """Returns u jet [C, B, o] and the stash the reverse pass needs."""
    xj = input_jet(cs, X)
    st = {"xj": xj}
    a = xj @ net.Win.T
    a[0] += net.bin
    s = act_fwd(cs, net.act, a)
    st["in"] = (a, s)
    st["layers"] = []
    if net.kind == KIND_MLP:
        for W, b in net.hidden:
            a = lin(cs, W, None, b, s, None)
            y = act_fwd(cs, net.act, a)
            st["layers"].append((s, a, y))
            s = y
    else:
        for g in net.gates:
            aZ = lin(cs, *g["Z"][0:1], g["Z"][1], g["Z"][2], s, xj)
            aG = lin(cs, g["G"][0], g["G"][1], g["G"][2], s, xj)
            aR = lin(cs, g["R"][0], g["R"][1], g["R"][2], s, xj)
            Z, G, R = (act_fwd(cs, net.gact, t) for t in (aZ, aG, aR))
            sR = prod_fwd(cs, s, R)
            aH = lin(cs, g["H"][0], g["H"][1], g["H"][2], sR, xj)
            Hh = act_fwd(cs, net.gact, aH)
            omG = -G
            omG[0] = 1 - G[0]
            s_new = prod_fwd(cs, omG, Hh) + prod_fwd(cs, Z, s)
            st["layers"].append(dict(s=s, aZ=aZ, aG=aG, aR=aR, aH=aH, Z=Z, G=G, R=R, H=Hh,
                                     sR=sR, omG=omG))
            s = s_new
    st["s_last"] = s
    u = s @ net.Wout.T
    u[0] += net.bout
    return u, st


def reverse(net: Net, cs: ChanSet, st, ubar):
    """ubar [C, B, o] -> flat gradient (same layout as theta)."""
    g = np.zeros_like(net.theta)
    gn = net.grad_views(g)
    xj = st["xj"]
    s = st["s_last"]
    gn.Wout[...] += np.einsum("cbo,cbh->oh", ubar, s)
    gn.bout[...] += ubar[0].sum(0)
    sbar = ubar @ net.Wout

    def lin_adj(gW, gU, gb, abar, s_in, has_u=True):
        """returns cotangent of s_in; accumulates weight grads (views may be
        transposed views into the flat gradient for DGM_RAW)."""
        gW += np.einsum("cbo,cbi->oi", abar, s_in)
        if has_u:
            gU += np.einsum("cbo,cbi->oi", abar, xj)
        gb += abar[0].sum(0)

    if net.kind == KIND_MLP:
        for li in reversed(range(net.L)):
            s_in, a, y = st["layers"][li]
            abar = act_adj(cs, net.act, sbar, y[0], a)
            W, b = net.hidden[li]
            gW, gb = gn.hidden[li]
            lin_adj(gW, None, gb, abar, s_in, has_u=False)
            sbar = abar @ W
    else:
        for li in reversed(range(net.L)):
            L = st["layers"][li]
            gt, gg = net.gates[li], gn.gates[li]
            Hbar = prod_adj(cs, sbar, L["omG"])
            Gbar = -prod_adj(cs, sbar, L["H"])
            Zbar = prod_adj(cs, sbar, L["s"])
            s_bar = prod_adj(cs, sbar, L["Z"])
            aHbar = act_adj(cs, net.gact, Hbar, L["H"][0], L["aH"])
            lin_adj(gg["H"][0], gg["H"][1], gg["H"][2], aHbar, L["sR"])
            sRbar = aHbar @ gt["H"][0]
            Rbar = prod_adj(cs, sRbar, L["s"])
            s_bar += prod_adj(cs, sRbar, L["R"])
            for nm, ybar, a in (("Z", Zbar, L["aZ"]), ("G", Gbar, L["aG"]), ("R", Rbar, L["aR"])):
                abar = act_adj(cs, net.gact, ybar, L[nm][0], a)
                lin_adj(gg[nm][0], gg[nm][1], gg[nm][2], abar, L["s"])
                s_bar += abar @ gt[nm][0]
            sbar = s_bar
    a, s0 = st["in"]
    abar = act_adj(cs, net.act, sbar, s0[0], a)
    gn.Win[...] += np.einsum("cbo,cbi->oi", abar, xj)
    gn.bin[...] += abar[0].sum(0)
    return g


# ---------------------------------------------------------------- problems
def _net(spec, theta):
    kind, d, o, H, L, act = (int(v) for v in spec)
    return Net(kind, d, o, H, L, act, theta)


def heat_step(spec, theta, X, X0, XBD1, XBD2, x_bd1, x_bd2, kappa=1.0):
    net = _net(spec, theta)
    B = X.shape[0]
    u, st = forward(net, CS_HEAT, X)
    r = u[2] - kappa * u[3]
    ubar = np.zeros_like(u)
    ubar[2] = 2 * r / B
    ubar[3] = -2 * kappa * r / B
    g = reverse(net, CS_HEAT, st, ubar)
    loss = (r ** 2).sum()
    for Xc, tgt in ((X0, np.sin(X0[:, 0:1])), (XBD1, x_bd1), (XBD2, x_bd2)):
        uc, stc = forward(net, CS_V, Xc)
        e = uc[0] - tgt
        loss += (e ** 2).sum()
        g += reverse(net, CS_V, stc, (2 * e / B)[None])
    return loss / B, g


def ode_step(spec, theta, t, t0, y_ic):
    net = _net(spec, theta)
    B = t.shape[0]
    u, st = forward(net, CS_D1O1, t)
    r = u[1] + u[0]
    ubar = np.stack([2 * r / B, 2 * r / B])
    g = reverse(net, CS_D1O1, st, ubar)
    u0, st0 = forward(net, CS_V, t0)
    e = u0[0] - y_ic
    g += reverse(net, CS_V, st0, (2 * e / B)[None])
    return ((r ** 2).sum() + (e ** 2).sum()) / B, g


def fhn_step(spec, theta, t, t0, y_ic, I=0.5, alpha=0.7, beta=0.8, tau=2.5):
    net = _net(spec, theta)
    B = t.shape[0]
    u, st = forward(net, CS_D1O1, t)
    Y, W, dY, dW = u[0, :, 0], u[0, :, 1], u[1, :, 0], u[1, :, 1]
    rx = dY + (Y ** 3 / 3.0 + W - I - Y)
    ry = dW + (beta * W - alpha - Y) / tau
    ubar = np.zeros_like(u)
    ubar[1, :, 0] = 2 * rx / B
    ubar[1, :, 1] = 2 * ry / B
    ubar[0, :, 0] = 2 * rx / B * (Y * Y - 1) - 2 * ry / B / tau
    ubar[0, :, 1] = 2 * rx / B + 2 * ry / B * beta / tau
    g = reverse(net, CS_D1O1, st, ubar)
    u0, st0 = forward(net, CS_V, t0)
    e = u0[0] - y_ic
    g += reverse(net, CS_V, st0, (2 * e / (2 * B))[None])
    return (rx ** 2).sum() / B + (ry ** 2).sum() / B + (e ** 2).sum() / (2 * B), g


def fredholm_step(spec, theta, x, T):
    """T [k, B, 1]: per-point MC nodes (fredholm.py:64-69)."""
    net = _net(spec, theta)
    k, B = T.shape[0], x.shape[0]
    dr = math.pi / (2 * k)
    ux, stx = forward(net, CS_V, x)
    nodes = T.reshape(k * B, 1)
    un, stn = forward(net, CS_V, nodes)
    w = (np.sin(x)[None] * np.cos(T)).reshape(k * B, 1)      # sin x cos t_j
    integral = (w * un[0]).reshape(k, B, 1).sum(0) * dr
    r = ux[0] - np.sin(x) - integral
    g = reverse(net, CS_V, stx, (2 * r / B)[None])
    rb = np.broadcast_to(-(2 * r / B) * dr, (k, B, 1)).reshape(k * B, 1)
    g += reverse(net, CS_V, stn, (rb * w)[None])
    return (r ** 2).sum() / B, g


def jets_full(spec, theta, X):
    """y [B,o], J [B,o,d], Hs [B,o,d,d]."""
    net = _net(spec, theta)
    d = net.d
    cs = CS_D1O2 if d == 1 else CS_D2O2
    u, _ = forward(net, cs, X)
    B = X.shape[0]
    y = u[0]
    J = np.stack([u[1 + k] for k in range(d)], -1)
    Hs = np.zeros((B, net.o, d, d), X.dtype)
    for q, (i, j) in enumerate(cs.pairs):
        Hs[:, :, i, j] = u[1 + d + q]
        Hs[:, :, j, i] = u[1 + d + q]
    return y, J, Hs
