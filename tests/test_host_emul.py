"""Pipeline orchestration (csrc/dgmk_pipeline.h, dgmk_capi_impl.h, dgmk_ops.h) compiled
with a plain-loop host backend (tests/host_emul, TEST ONLY) against the executed-
reference golden vectors.  Proves carving / ordering / packing / scaling before the
GPU run; the CUDA kernels themselves are covered by the -m gpu tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden, golden_names, rel
import importlib.util

_spec = importlib.util.spec_from_file_location(
    "_cabi_for_emul", os.path.join(ROOT, "differential_equations_dnn_b200", "_cabi.py"))
cabi = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(cabi)

EMUL = os.path.join(ROOT, "tests", "host_emul", "libdgmk_emul.so")


@pytest.fixture(scope="module")
def lib():
    subprocess.run([os.path.join(ROOT, "tests", "host_emul", "build.sh")], check=True)
    return cabi.bind(C.CDLL(EMUL))


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def desc_of(g):
    k, d, o, H, L, act = (int(v) for v in g["spec"])
    return cabi.make_desc(k, d, o, H, L, act)


def run_step(lib, prob, g, ws_bytes=None, Bg=None):
    desc = desc_of(g)
    theta = f32(g["theta"])
    grad = np.full_like(theta, np.nan)
    loss = np.zeros(1, np.float32)
    if prob == "heat":
        B = g["X"].shape[0]
        cls, k = cabi.WS_HEAT, 0
    elif prob in ("ode", "fhn"):
        B = g["t"].shape[0]
        cls, k = (cabi.WS_ODE if prob == "ode" else cabi.WS_FHN), 0
    else:
        B = g["x"].shape[0]
        cls, k = cabi.WS_FREDHOLM, g["T"].shape[0]
    Bg = Bg or B
    need = lib.dgmk_workspace_bytes(C.byref(desc), cls, B, k)
    assert need > 0
    wsb = ws_bytes or need
    ws = np.zeros(wsb // 4 + 16, np.float32)
    if prob == "heat":
        a = [f32(g[n]) for n in ("X", "X0", "XBD1", "XBD2", "x_bd1", "x_bd2")]
        rc = lib.dgmk_heat_step(C.byref(desc), P(theta), *[P(z) for z in a], B, Bg, 1.0, P(loss), P(grad),
                                P(ws), wsb, None)
    elif prob in ("ode", "fhn"):
        a = [f32(g[n]) for n in ("t", "t0", "y_ic")]
        fn = lib.dgmk_ode_step if prob == "ode" else lib.dgmk_fhn_step
        rc = fn(C.byref(desc), P(theta), *[P(z) for z in a], B, Bg, P(loss), P(grad), P(ws), wsb, None)
    else:
        x, T = f32(g["x"]), f32(g["T"])
        rc = lib.dgmk_fredholm_step(C.byref(desc), P(theta), P(x), P(T), B, k, Bg, P(loss), P(grad), P(ws),
                                    wsb, None)
    assert rc == 0, lib.dgmk_last_error()
    return float(loss[0]), grad


def entry_slices(lib, desc):
    n = lib.dgmk_param_layout(C.byref(desc), -1, None, None, None, None)
    out = []
    for i in range(n):
        off, r, c, lv = C.c_int64(), C.c_int32(), C.c_int32(), C.c_int32()
        assert lib.dgmk_param_layout(C.byref(desc), i, C.byref(off), C.byref(r), C.byref(c), C.byref(lv)) == 0
        out.append((off.value, r.value * max(c.value, 1), lv.value))
    return out


def check(lib, g, loss, grad, tol=1e-5):
    assert abs(loss - float(g["loss"])) <= tol * abs(float(g["loss"])), (loss, float(g["loss"]))
    desc = desc_of(g)
    assert lib.dgmk_param_count(C.byref(desc)) == g["theta"].size
    worst = 0.0
    for off, n, live in entry_slices(lib, desc):
        ref = g["grad"][off:off + n]
        mine = grad[off:off + n]
        assert np.all(np.isfinite(mine))
        if not live:
            assert np.all(mine == 0) and not g["live"][off:off + n].any()
            continue
        if np.linalg.norm(ref) == 0:
            assert np.linalg.norm(mine) < 1e-7
        else:
            worst = max(worst, rel(mine, ref))
    assert worst < tol, worst


PROBS = ("heat", "ode", "fhn", "fredholm")
CASES = [(p, n) for p in PROBS for n in golden_names(p + "_") if "driver" not in n]


@pytest.mark.parametrize("prob,name", CASES)
def test_step_matches_reference(lib, prob, name):
    g = golden(name)
    loss, grad = run_step(lib, prob, g)
    check(lib, g, loss, grad)


@pytest.mark.parametrize("prob,name", [c for c in CASES if c[0] != "fredholm"])
def test_inplace_reverse_matches_reference(lib, prob, name):
    """The reverse pass of the resident-tile step (csrc/dgmk_tile.cuh) overwrites the forward stash as it consumes
    it (pre-activation cotangents over the a-form gates, (s*R)bar over s*R, one state-cotangent buffer): the same
    orchestration template in that mode, host loops as the backend, against the executed reference -- and
    bit-identical to the out-of-place mode."""
    g = golden(name)
    loss0, grad0 = run_step(lib, prob, g)
    lib.dgmk_emul_set_inplace(1)
    try:
        loss, grad = run_step(lib, prob, g)
    finally:
        lib.dgmk_emul_set_inplace(0)
    check(lib, g, loss, grad)
    assert loss == loss0 and np.array_equal(grad, grad0)


@pytest.mark.parametrize("name", golden_names("fredholm_"))
@pytest.mark.parametrize("points,nodes,inplace", [(8, 3, 1), (5, 64, 1), (1, 1, 0), (32, 2, 0)])
def test_fredholm_blocks_match_reference(lib, name, points, nodes, inplace):
    """The Fredholm body the resident-tile kernel runs (dgmk_steps.h fredholm_block: blocks of points, their k nodes
    walked in sub-tiles, node rows evaluated twice when they span several sub-tiles) with host loops as the backend:
    ragged blocks, sub-tiles that do and do not divide k, one sub-tile holding all k nodes."""
    g = golden(name)
    lib.dgmk_emul_set_fredholm_blocks(points, nodes)
    lib.dgmk_emul_set_inplace(inplace)
    try:
        loss, grad = run_step(lib, "fredholm", g)
    finally:
        lib.dgmk_emul_set_fredholm_blocks(0, 0)
        lib.dgmk_emul_set_inplace(0)
    check(lib, g, loss, grad)


@pytest.mark.parametrize("prob,name", [("heat", "heat_dgm_h32l1"), ("fredholm", "fredholm_dgmraw_h32l1_k7"),
                                       ("fhn", "fhn_dgm_h64l2")])
def test_chunked_equals_unchunked(lib, prob, name):
    """A workspace too small for the batch forces several chunks; results must agree."""
    g = golden(name)
    desc = desc_of(g)
    cls = {"heat": cabi.WS_HEAT, "fhn": cabi.WS_FHN, "fredholm": cabi.WS_FREDHOLM}[prob]
    k = g["T"].shape[0] if prob == "fredholm" else 0
    small = lib.dgmk_workspace_bytes(C.byref(desc), cls, 7, k)
    loss, grad = run_step(lib, prob, g, ws_bytes=small)
    check(lib, g, loss, grad)


def test_data_parallel_shards_sum_to_global(lib):
    """each rank scales by 1/B_global; SUM over ranks == single-process result."""
    g = golden("heat_dgm_h32l1")
    B = g["X"].shape[0]
    tot_l, tot_g = 0.0, 0.0
    for lo, hi in ((0, 40), (40, B)):
        sub = dict(g)
        for n in ("X", "X0", "XBD1", "XBD2", "x_bd1", "x_bd2"):
            sub[n] = g[n][lo:hi]
        l, gr = run_step(lib, "heat", sub, Bg=B)
        tot_l, tot_g = tot_l + l, tot_g + gr
    check(lib, g, tot_l, tot_g)


def test_errors(lib):
    g = golden("ode_mlp_relu_h32l1")
    desc = desc_of(g)
    bad = cabi.make_desc(0, 3, 1, 32, 1, 0)
    assert lib.dgmk_param_count(C.byref(bad)) == -1
    assert b"input_dim" in lib.dgmk_last_error()
    theta = f32(g["theta"])
    ws = np.zeros(64, np.float32)
    loss = np.zeros(1, np.float32)
    a = [f32(g[n]) for n in ("t", "t0", "y_ic")]
    rc = lib.dgmk_ode_step(C.byref(desc), P(theta), *[P(z) for z in a], 64, 64, P(loss), P(theta.copy()), P(ws), 256, None)
    assert rc == -2  # DGMK_EWORKSPACE
    rc = lib.dgmk_heat_step(C.byref(desc), P(theta), *[P(theta)] * 6, 64, 64, 1.0, P(loss), P(theta), P(ws), 256, None)
    assert rc == -1  # wrong dims for heat


@pytest.mark.parametrize("cs,d", [(0, 1), (0, 2), (1, 1), (2, 2), (3, 2), (4, 1), (5, 2)])
def test_input_map_adj_equals_abar_t_e(lib, cs, d):
    """grad[U | b] formed from the point coordinates (dgmk_ops.h::input_map_adj: the structural zeros and ones
    of the E rows folded in -- what the fused CUDA path accumulates next to the cotangents) equals Abar^T E with
    the E rows ExtInputFn writes, for every channel set."""
    lib.dgmk_emul_check_input_map_adj.restype = C.c_double
    lib.dgmk_emul_check_input_map_adj.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint]
    for seed in (1, 2):
        assert lib.dgmk_emul_check_input_map_adj(cs, 300, d, seed) < 2e-6


@pytest.mark.parametrize("cs", [0, 1, 2])
@pytest.mark.parametrize("tanh_gates", [1, 0])
@pytest.mark.parametrize("V", [2, 4])
def test_rev1_units_per_thread_variant_is_bit_identical(lib, cs, tanh_gates, V):
    """DgmRev1Fn::runv<V> (V units per call: the CUDA backend's rev1_ev_kernel) stores exactly what the
    one-unit functor stores, and hands its sink exactly those cotangents."""
    lib.dgmk_emul_check_rev1_runv.restype = C.c_int
    lib.dgmk_emul_check_rev1_runv.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint]
    assert lib.dgmk_emul_check_rev1_runv(cs, tanh_gates, V, 37, 5) == 0


@pytest.mark.parametrize("n,lo,hi,seed,stream,step_dev,step_add", [
    (1, 0.0, 1.0, 0, 0, None, 0), (7, 0.0, 1.01, 1234, 0, None, 5), (64, -2.0, 3.5, 2 ** 63 + 12345, 9, 41, 1),
    (1003, 0.0, np.pi / 2, 2 ** 40 + 3, 257, 2 ** 33 + 5, 0)])
def test_philox_uniform_functor_is_bit_exact(lib, n, lo, hi, seed, stream, step_dev, step_add):
    """PhiloxUniformFn (csrc/dgmk_ops.h, the code dgmk_sample_uniform launches) compiled for the host against
    oracle/philox_np.py: ragged tails, non-zero lo, 64-bit seeds, the step read through the device pointer."""
    from oracle import philox_np as PH
    out = np.full(n + 3, np.nan, np.float32)
    sd = np.array([step_dev if step_dev is not None else 0], np.int64)
    rc = lib.dgmk_sample_uniform(P(out), n, lo, hi, seed, stream, P(sd) if step_dev is not None else None, step_add, None)
    assert rc == 0
    want = PH.uniform(n, lo, hi, seed, stream, (step_dev or 0) + step_add)
    assert np.array_equal(out[:n], want)
    assert np.all(np.isnan(out[n:]))   # nothing written past n


@pytest.mark.parametrize("B", [1, 6, 64, 203])
def test_philox_heat_functor_is_bit_exact(lib, B):
    from oracle import philox_np as PH
    bufs = [np.full((B + 1, 2), np.nan, np.float32) for _ in range(4)]
    sd = np.array([17], np.int64)
    rc = lib.dgmk_sample_heat(P(bufs[0]), P(bufs[1]), P(bufs[2]), P(bufs[3]), B, np.pi, 3.0, np.pi, 99, P(sd), 2, None)
    assert rc == 0
    for got, want in zip(bufs, PH.heat(B, np.pi, 3.0, np.pi, 99, 19)):
        assert np.array_equal(got[:B], want)
        assert np.all(np.isnan(got[B:]))
    X, X0, B1, B2 = (b[:B] for b in bufs)
    assert X[:, 0].max() < np.float32(np.pi) + 1e-6 and X[:, 1].max() < 3.0 and X.min() >= 0.0
    assert np.array_equal(X0[:, 0], X[:, 0]) and np.array_equal(B1[:, 1], X[:, 1]) and np.all(B2[:, 0] == np.float32(np.pi))
