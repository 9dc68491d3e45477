"""Flat-parameter plumbing shared by the drop-in network classes.

The kernels, the fused Adam update and the data-parallel all-reduce all want ONE
pointer, so every nn.Parameter of a network is a view into one contiguous FP32
buffer laid out in `named_parameters()` order -- the order of the reference modules
(neural_networks.py:184-228, :134-160; dgm_net.py:75-101), which is also the order
the C ABI's `dgmk_param_layout` reports.  state_dict keys / shapes are unchanged, so
checkpoints round-trip with the reference classes (checkpoint row).
"""
from __future__ import annotations

import weakref

import torch
from torch import nn

from . import _cabi, kernels


class DeferredOutput:
    """What `net(x)` returns inside `deferred_forward(net)`: a record of the call.

    The reference drivers call `y = net(t); y0 = net(t0); loss = dgm_loss_func(y, y0,
    t, y_ic)` (simple_ode.py:98-101).  Evaluating `net(t)` eagerly and then throwing
    the result away when the fused step kernel recomputes it would waste ~20 % of the
    step, so our drivers record the calls and let `dgm_loss_func` launch one fused
    step.  Only the package's own loss functions accept this object.
    """

    __slots__ = ("net", "x")

    def __init__(self, net, x):
        self.net, self.x = net, x


class deferred_forward:
    def __init__(self, net):
        self.net = net

    def __enter__(self):
        self.net._deferred = True
        return self.net

    def __exit__(self, *exc):
        self.net._deferred = False
        return False


#: flat buffer address -> owning module (weak): how FusedAdam finds the network from `net.parameters()`.
#: Nothing is attached to the Parameters themselves, so modules pickle / deepcopy like the reference's.
_OWNERS: "weakref.WeakValueDictionary[int, FlatParamModule]" = weakref.WeakValueDictionary()


def owner_of(params):
    """The FlatParamModule whose parameter list is exactly `params` (None if there is none)."""
    if not params:
        return None
    net = _OWNERS.get(params[0].data_ptr())
    if net is None or len(net._plist) != len(params) or any(a is not b for a, b in zip(net._plist, params)):
        return None
    return net


class FlatParamModule(nn.Module):
    """Base of MLP / DGM: flat storage + kernel-backed forward."""

    #: highest input-derivative order `forward` prepares when `x.requires_grad`
    #: (2 = value, Jacobian and Hessian: enough for every reference loss).
    jet_order = 2

    def _finish_init(self, kind, d, o, H, L, act):
        self._desc_args = (kind, d, o, H, L, act)
        self._deferred = False
        self._flat = None
        self._flatten()
        self._check_layout()

    # ---- flat storage -------------------------------------------------------------
    @property
    def desc(self):
        return _cabi.make_desc(*self._desc_args)

    def _flatten(self):
        params = list(self.parameters())
        flat = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
        off = 0
        for p in params:
            n = p.numel()
            p.data = flat[off:off + n].view(p.shape)
            off += n
        object.__setattr__(self, "_flat", flat)
        object.__setattr__(self, "_plist", params)
        _OWNERS[flat.data_ptr()] = self

    def __setstate__(self, state):
        # pickle / torch.save(net) / copy.deepcopy(net) restore the parameters one by one: re-tie them to one
        # buffer and register the copy
        super().__setstate__(state)
        self._flatten()

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        self._flatten()  # .to()/.cuda()/.float() re-allocate per parameter: re-tie them
        return out

    def flat_theta(self):
        """The single FP32 buffer every parameter is a view of."""
        ps = self._plist
        f = self._flat
        if ps[0].data_ptr() != f.data_ptr() or \
                ps[-1].data_ptr() != f.data_ptr() + 4 * (f.numel() - ps[-1].numel()):
            self._flatten()
            f = self._flat
        return f

    def param_slices(self):
        """[(param, offset, numel, live)]; `live` False = never receives a gradient."""
        out, off = [], 0
        dead = getattr(self, "_dead_prefixes", ())
        for name, p in self.named_parameters():
            out.append((p, off, p.numel(), not name.startswith(dead) if dead else True))
            off += p.numel()
        return out

    def live_mask(self):
        f = self.flat_theta()
        m = torch.ones(f.numel(), dtype=torch.uint8)
        for _, off, n, live in self.param_slices():
            if not live:
                m[off:off + n] = 0
        return m.to(f.device)

    def _check_layout(self):
        """Python layout == C ABI layout (only when the library is built)."""
        import os
        if not os.path.exists(_cabi.LIB_PATH):
            return
        lay = kernels.param_layout(self.desc)
        mine = [(off, tuple(p.shape), live) for p, off, _, live in self.param_slices()]
        assert len(lay) == len(mine), "parameter list differs from the C ABI layout"
        for (off, r, c, live), (moff, shape, mlive) in zip(lay, mine):
            cshape = (r,) if c == 0 else (r, c)
            assert off == moff and cshape == shape and live == mlive, "layout mismatch with C ABI"

    # ---- forward --------------------------------------------------------------------
    def forward(self, x):
        if self._deferred:
            return DeferredOutput(self, x)
        from . import autograd as ag
        squeeze = x.dim() == 1  # gridEvaluation passes a 1-D [d] point (simple_ode.py:128-131)
        x2 = x.reshape(1, -1) if squeeze else x
        if x2.dim() != 2 or x2.shape[1] != self._desc_args[1]:
            raise ValueError(f"expected input [*, {self._desc_args[1]}], got {tuple(x.shape)}")
        y = ag.module_forward(self, x2)
        return y.reshape(-1) if squeeze else y
