"""Thin torch-tensor front end of the C ABI (include/dgmk.h): pointer plumbing only.

PyTorch is used for device memory, streams and (elsewhere) torch.distributed; all
arithmetic happens in csrc/libdgmk.so.  Every function here rejects CPU tensors --
there is no fallback path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi
from ._cabi import NetDesc, DgmkError  # noqa: F401

_WS_CACHE: dict = {}
# cap on the scratch a single call may hold; the steps chunk the batch to fit.
import os as _os
WORKSPACE_CAP_BYTES = int(float(_os.environ.get("DGMK_WORKSPACE_GB", "48")) * (1 << 30))


def _dev_f32(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise DgmkError("dgmk kernels need CUDA tensors (no CPU fallback exists)")
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise DgmkError("dgmk kernels need contiguous float32 tensors")


def _expect(t, shape, name):
    """The kernels index every row of every operand: a shorter or broadcastable companion would be read out of
    bounds, so shapes are checked here (leading dimension exact, element count exact)."""
    want = 1
    for v in shape:
        want *= v
    if t.dim() == 0 or t.shape[0] != shape[0] or t.numel() != want:
        raise DgmkError(f"{name}: expected shape {list(shape)}, got {list(t.shape)}")


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def workspace_bytes(desc, ws_class, B, k=0):
    lib = _cabi.load()
    n = lib.dgmk_workspace_bytes(C.byref(desc), ws_class, B, k)
    if n == 0:
        raise DgmkError(lib.dgmk_last_error().decode())
    return n


def get_workspace(device, nbytes):
    """One cached scratch buffer per (device, stream); grows on demand."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _WS_CACHE.get(key)
    if buf is None or buf.numel() < nbytes:
        _WS_CACHE[key] = None
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _WS_CACHE[key] = buf
    return buf


def release_workspaces():
    _WS_CACHE.clear()


def param_count(desc):
    lib = _cabi.load()
    n = lib.dgmk_param_count(C.byref(desc))
    if n < 0:
        raise DgmkError(lib.dgmk_last_error().decode())
    return n


def param_layout(desc, lib=None):
    """[(offset, rows, cols, live)] in named_parameters() order (cols == 0: 1-D)."""
    lib = lib or _cabi.load()
    n = lib.dgmk_param_layout(C.byref(desc), -1, None, None, None, None)
    if n < 0:
        raise DgmkError(lib.dgmk_last_error().decode())
    out = []
    for i in range(n):
        off, r, c, lv = C.c_int64(), C.c_int32(), C.c_int32(), C.c_int32()
        _cabi.check(lib.dgmk_param_layout(C.byref(desc), i, C.byref(off), C.byref(r), C.byref(c),
                                          C.byref(lv)), lib)
        out.append((off.value, r.value, c.value, bool(lv.value)))
    return out


_WS_BYTES: dict = {}


def _step_common(desc, ws_class, theta, B, k, ws):
    if ws is None:
        key = (desc.kind, desc.input_dim, desc.output_dim, desc.hidden_size, desc.num_layers, desc.activation,
               ws_class, B, k)
        nbytes = _WS_BYTES.get(key)   # the small-batch loops call this every step: one C call per shape, not per step
        if nbytes is None:
            nbytes = _WS_BYTES[key] = workspace_bytes(desc, ws_class, B, k)
        ws = get_workspace(theta.device, min(nbytes, WORKSPACE_CAP_BYTES))
    out = torch.empty(theta.numel() + 1, dtype=torch.float32, device=theta.device)
    return ws, out


def heat_step(desc, theta, x, x0, xbd1, xbd2, x_bd1, x_bd2, kappa=1.0, B_global=None, ws=None):
    """heat.py:50-95 + loss.backward(): returns a [P+1] tensor = [grad_theta | loss]."""
    _dev_f32(theta, x, x0, xbd1, xbd2, x_bd1, x_bd2)
    lib = _cabi.load()
    B, d = x.shape[0], desc.input_dim
    for t, nm in ((x, "x"), (x0, "x0"), (xbd1, "xbd1"), (xbd2, "xbd2")):
        _expect(t, (B, d), nm)
    _expect(x_bd1, (B, 1), "x_bd1")
    _expect(x_bd2, (B, 1), "x_bd2")
    ws, out = _step_common(desc, _cabi.WS_HEAT, theta, B, 0, ws)
    P = theta.numel()
    with torch.cuda.device(theta.device):
        rc = lib.dgmk_heat_step(C.byref(desc), _ptr(theta), _ptr(x), _ptr(x0), _ptr(xbd1), _ptr(xbd2),
                                _ptr(x_bd1), _ptr(x_bd2), B, B_global or B, float(kappa),
                                C.c_void_p(out.data_ptr() + 4 * P), _ptr(out), _ptr(ws), ws.numel(),
                                _stream(theta.device))
    _cabi.check(rc, lib)
    return out


def _ode_like(fn_name, ws_class, desc, theta, t, t0, y_ic, B_global, ws):
    _dev_f32(theta, t, t0, y_ic)
    lib = _cabi.load()
    B = t.shape[0]
    _expect(t, (B, 1), "t")
    _expect(t0, (B, 1), "t0")
    _expect(y_ic, (B, desc.output_dim), "y_ic")
    ws, out = _step_common(desc, ws_class, theta, B, 0, ws)
    P = theta.numel()
    with torch.cuda.device(theta.device):
        rc = getattr(lib, fn_name)(C.byref(desc), _ptr(theta), _ptr(t), _ptr(t0), _ptr(y_ic), B,
                                   B_global or B, C.c_void_p(out.data_ptr() + 4 * P), _ptr(out),
                                   _ptr(ws), ws.numel(), _stream(theta.device))
    _cabi.check(rc, lib)
    return out


def ode_step(desc, theta, t, t0, y_ic, B_global=None, ws=None):
    """simple_ode.py:41-63 (+ the two net calls of the driver, :98-99) + backward."""
    return _ode_like("dgmk_ode_step", _cabi.WS_ODE, desc, theta, t, t0, y_ic, B_global, ws)


def fhn_step(desc, theta, t, t0, y_ic, B_global=None, ws=None):
    """fitzhugh_nagumo.py:53-97 (+ driver :137-138) + backward."""
    return _ode_like("dgmk_fhn_step", _cabi.WS_FHN, desc, theta, t, t0, y_ic, B_global, ws)


def fredholm_step(desc, theta, x, nodes, B_global=None, ws=None):
    """fredholm.py:47-74 + backward; nodes [k,B,1] = the k rand_like draws in order."""
    _dev_f32(theta, x, nodes)
    lib = _cabi.load()
    B, k = x.shape[0], nodes.shape[0]
    _expect(x, (B, 1), "x")
    _expect(nodes, (k, B, 1), "nodes")
    ws, out = _step_common(desc, _cabi.WS_FREDHOLM, theta, B, k, ws)
    P = theta.numel()
    with torch.cuda.device(theta.device):
        rc = lib.dgmk_fredholm_step(C.byref(desc), _ptr(theta), _ptr(x), _ptr(nodes), B, k, B_global or B,
                                    C.c_void_p(out.data_ptr() + 4 * P), _ptr(out), _ptr(ws), ws.numel(),
                                    _stream(theta.device))
    _cabi.check(rc, lib)
    return out


def jet_forward(desc, theta, x, order):
    """net(x) with input derivatives: returns (Y, J, Hs, stash workspace)."""
    _dev_f32(theta, x)
    lib = _cabi.load()
    B, d, o = x.shape[0], desc.input_dim, desc.output_dim
    nbytes = workspace_bytes(desc, _cabi.WS_JET0 + order, B, 0)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)  # stash: owned by the autograd node
    Y = torch.empty(B, o, device=x.device)
    J = torch.empty(B, o, d, device=x.device) if order >= 1 else None
    Hs = torch.empty(B, o, d, d, device=x.device) if order >= 2 else None
    with torch.cuda.device(x.device):
        rc = lib.dgmk_jet_forward(C.byref(desc), _ptr(theta), _ptr(x), B, order, _ptr(Y), _ptr(J), _ptr(Hs),
                                  _ptr(ws), ws.numel(), _stream(x.device))
    _cabi.check(rc, lib)
    return Y, J, Hs, ws


def jet_reverse(desc, theta, x, order, gY, gJ, gHs, ws):
    _dev_f32(theta, x, gY, gJ, gHs)
    lib = _cabi.load()
    grad = torch.empty_like(theta)
    with torch.cuda.device(x.device):
        rc = lib.dgmk_jet_reverse(C.byref(desc), _ptr(theta), _ptr(x), x.shape[0], order, _ptr(gY), _ptr(gJ),
                                  _ptr(gHs), _ptr(grad), _ptr(ws), ws.numel(), _stream(x.device))
    _cabi.check(rc, lib)
    return grad


def evaluate(desc, theta, x, ws=None):
    """Value-only batched forward (replaces the point-by-point gridEvaluation loops)."""
    _dev_f32(theta, x)
    lib = _cabi.load()
    B = x.shape[0]
    if ws is None:
        nbytes = workspace_bytes(desc, _cabi.WS_JET0, min(B, 1 << 16), 0)
        ws = get_workspace(x.device, nbytes)
    Y = torch.empty(B, desc.output_dim, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.dgmk_eval(C.byref(desc), _ptr(theta), _ptr(x), B, _ptr(Y), _ptr(ws), ws.numel(),
                           _stream(x.device))
    _cabi.check(rc, lib)
    return Y


def adam_step(theta, m, v, grad, live, lr, beta1, beta2, eps, step):
    """torch.optim.Adam defaults on flat buffers, one launch."""
    _dev_f32(theta, m, v, grad)
    lib = _cabi.load()
    with torch.cuda.device(theta.device):
        rc = lib.dgmk_adam(_ptr(theta), _ptr(m), _ptr(v), _ptr(grad), _ptr(live), theta.numel(), float(lr),
                           float(beta1), float(beta2), float(eps), int(step), _stream(theta.device))
    _cabi.check(rc, lib)


def adam_step_dev(theta, m, v, grad, live, lr, beta1, beta2, eps, state):
    """Same update with the step counter in device memory (`state`: 2 x int64, zero-initialised):
    no host-side state, so a captured CUDA graph can replay it."""
    _dev_f32(theta, m, v, grad)
    if not (state.is_cuda and state.dtype == torch.int64 and state.numel() >= 2):
        raise DgmkError("adam_step_dev needs a CUDA int64 state tensor of 2 elements")
    lib = _cabi.load()
    with torch.cuda.device(theta.device):
        rc = lib.dgmk_adam_dev(_ptr(theta), _ptr(m), _ptr(v), _ptr(grad), _ptr(live), theta.numel(), float(lr),
                               float(beta1), float(beta2), float(eps), _ptr(state), _stream(theta.device))
    _cabi.check(rc, lib)


def _step_ptr(step):
    if step is None:
        return None
    if not (step.is_cuda and step.dtype == torch.int64 and step.numel() >= 1):
        raise DgmkError("the sampler's step counter must be a CUDA int64 tensor")
    return _ptr(step)


def sample_uniform(out, lo, hi, seed, stream_id=0, step=None, step_add=0):
    """out[...] = lo + (hi - lo) * u, u ~ U[0, 1) from Philox4x32-10 (include/dgmk.h: dgmk_sample_uniform; bit-exact oracle
    oracle/philox_np.py).  `step`: device int64 counter read by the kernel (CUDA-graph replays), plus `step_add`."""
    _dev_f32(out)
    lib = _cabi.load()
    with torch.cuda.device(out.device):
        rc = lib.dgmk_sample_uniform(_ptr(out), out.numel(), float(lo), float(hi), int(seed) & (2 ** 64 - 1), int(stream_id),
                                     _step_ptr(step), int(step_add), _stream(out.device))
    _cabi.check(rc, lib)
    return out


def sample_heat(X, X0, XBD1, XBD2, xmax, tmax, xbd2, seed, step=None, step_add=0):
    """The four [B, 2] operand blocks of a heat step (heat.py:125-134) in one launch (dgmk_sample_heat)."""
    _dev_f32(X, X0, XBD1, XBD2)
    B = X.shape[0]
    for t in (X, X0, XBD1, XBD2):
        if tuple(t.shape) != (B, 2):
            raise DgmkError("sample_heat needs four [B, 2] tensors")
    lib = _cabi.load()
    with torch.cuda.device(X.device):
        rc = lib.dgmk_sample_heat(_ptr(X), _ptr(X0), _ptr(XBD1), _ptr(XBD2), B, float(xmax), float(tmax), float(xbd2),
                                  int(seed) & (2 ** 64 - 1), _step_ptr(step), int(step_add), _stream(X.device))
    _cabi.check(rc, lib)
