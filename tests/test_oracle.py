"""The oracle (both tiers) against the executed-reference golden vectors."""
import numpy as np
import pytest
import torch

from conftest import golden, golden_names, rel
from oracle import jets_np as jn
from oracle import ref_port as rp


def _spec(g):
    return rp.NetSpec(*[int(v) for v in g["spec"]])


def _entry_slices(spec):
    off, out = 0, []
    for name, shape, live in spec.entries():
        n = int(np.prod(shape))
        out.append((name, slice(off, off + n), live))
        off += n
    return out


def check_grad(spec, g_mine, g_ref, tol, live=None):
    for name, sl, _ in _entry_slices(spec):
        nr = np.linalg.norm(g_ref[sl])
        if nr == 0:
            assert np.linalg.norm(g_mine[sl]) <= 1e-12 + tol, name
        else:
            assert rel(g_mine[sl], g_ref[sl]) < tol, (name, rel(g_mine[sl], g_ref[sl]))


PROBLEMS = {
    "heat": (rp.heat_loss, jn.heat_step, ("X", "X0", "XBD1", "XBD2", "x_bd1", "x_bd2")),
    "ode": (rp.ode_loss, jn.ode_step, ("t", "t0", "y_ic")),
    "fhn": (rp.fhn_loss, jn.fhn_step, ("t", "t0", "y_ic")),
    "fredholm": (rp.fredholm_loss, jn.fredholm_step, ("x", "T")),
}
CASES = [(p, n) for p in PROBLEMS for n in golden_names(p + "_") if "driver" not in n]


@pytest.mark.parametrize("prob,name", CASES)
def test_ref_port_fp32_matches_reference(prob, name):
    g = golden(name)
    spec = _spec(g)
    fn, _, keys = PROBLEMS[prob]
    loss, grad = rp.loss_and_grad(fn, spec, torch.from_numpy(g["theta"]),
                                  *[torch.from_numpy(g[k]) for k in keys])
    assert abs(float(loss) - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    check_grad(spec, grad.numpy(), g["grad"], 1e-5)
    # dead parameters of neural_networks.DGM get exactly zero
    assert np.all(grad.numpy()[~g["live"]] == 0)


@pytest.mark.parametrize("prob,name", CASES)
def test_jets_fp64_matches_reference_fp64(prob, name):
    g = golden(name)
    _, fn, keys = PROBLEMS[prob]
    loss, grad = fn(g["spec"], g["theta"].astype(np.float64),
                    *[g[k].astype(np.float64) for k in keys])
    assert abs(loss - float(g["loss_f64"])) <= 1e-12 * abs(float(g["loss_f64"]))
    if "grad_f64" in g:
        check_grad(_spec(g), grad, g["grad_f64"], 1e-10)
    else:
        check_grad(_spec(g), grad, g["grad"], 2e-6)


@pytest.mark.parametrize("name", golden_names("jets_"))
def test_full_jets(name):
    g = golden(name)
    y, J, Hs = jn.jets_full(g["spec"], g["theta"].astype(np.float64), g["X"])
    assert rel(y, g["y"]) < 1e-12 and rel(J, g["J"]) < 1e-12 and rel(Hs, g["Hs"]) < 1e-11
    spec = _spec(g)
    y2, J2, H2 = rp.jets_full(spec, torch.from_numpy(g["theta"]).double(), torch.from_numpy(g["X"]))
    assert rel(y2, g["y"]) < 1e-12 and rel(J2, g["J"]) < 1e-12 and rel(H2, g["Hs"]) < 1e-11


def test_adam_matches_torch():
    g = golden("adam_5steps")
    th = torch.from_numpy(g["thetas"][0])
    m = torch.zeros_like(th)
    v = torch.zeros_like(th)
    live = torch.from_numpy(g["live"].astype(bool))
    for s in range(5):
        th, m, v = rp.adam_step(th, m, v, torch.from_numpy(g["grads"][s]), s + 1,
                                lr=float(g["lr"]), live=live)
        assert rel(th.numpy(), g["thetas"][s + 1]) < 1e-6
    assert rel(m.numpy(), g["m"]) < 1e-6 and rel(v.numpy(), g["v"]) < 1e-6


def test_known_answers():
    """Closed-form KATs: analytic solutions have zero residual."""
    x = np.linspace(0.1, 3.0, 7)
    t = np.linspace(0.0, 2.0, 7)
    u = np.sin(x) * np.exp(-t)
    assert np.allclose(-u - (-u), 0)  # u_t = -u = u_xx for sin(x)exp(-t), heat.py:36-47
    # Fredholm: y=2 sin x solves y = sin x + int_0^{pi/2} sin x cos t y(t) dt  (fredholm.py:40-44)
    tt = np.linspace(0, np.pi / 2, 200001)
    integral = np.trapezoid(np.cos(tt) * 2 * np.sin(tt), tt)
    assert abs(integral - 1.0) < 1e-9


def test_philox_kat():
    """oracle/philox_np.py against the known-answer vectors of Random123 (kat_vectors, philox4x32-10): the pin of the
    on-device sampler's generator (the CUDA kernel is compared with this oracle bit for bit in the -m gpu tests, the
    same functor compiled for the host in tests/test_host_emul.py)."""
    from oracle import philox_np as P
    kats = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kats:
        got = P.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert [int(v) for v in got] == want
    # stream layout: element i = word i % 4 of block i // 4; streams and steps are independent; u in [0, 1)
    w = P.words(10, seed=0x0123456789abcdef, stream_id=3, step=7)
    blk = P.philox4x32_10(np.array([[0, 0, 7, 3], [1, 0, 7, 3], [2, 0, 7, 3]], dtype=np.uint32),
                          np.array([[0x89abcdef, 0x01234567]], dtype=np.uint32)).reshape(-1)
    assert np.array_equal(w, blk[:10])
    u = P.uniform(100000, 0.0, 1.0, seed=5, stream_id=0, step=0)
    assert u.dtype == np.float32 and u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 5e-3 and abs(u.var() - 1 / 12) < 2e-3
    assert not np.array_equal(u[:16], P.uniform(16, 0.0, 1.0, seed=5, stream_id=1, step=0))
    assert not np.array_equal(u[:16], P.uniform(16, 0.0, 1.0, seed=5, stream_id=0, step=1))
