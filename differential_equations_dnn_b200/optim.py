"""Fused Adam on the flat parameter buffer (replaces torch.optim.Adam, heat.py:115,141).

Same defaults and arithmetic as torch.optim.Adam (betas 0.9/0.999, eps 1e-8, no weight
decay, no amsgrad); one kernel over [theta, m, v, grad] instead of ~12 tiny launches
per parameter tensor.  Parameters whose `.grad` is None are skipped, like
torch does -- that is how neural_networks.DGM's dead `dgm1` stays untouched.

`capturable=True` keeps the step counter in device memory (torch.optim.Adam's flag of the same
name), so that `step()` can be captured in a CUDA graph (heat.minimize_loss_dgm(cuda_graph=True)).
"""
from __future__ import annotations

import torch

from . import kernels


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, capturable=False):
        params = list(params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.capturable = bool(capturable)
        self._state = None
        from ._flat import owner_of
        self.net = owner_of(params)
        if self.net is None:
            raise ValueError("FusedAdam needs net.parameters() of a differential_equations_dnn_b200 network")
        self._m = self._v = self._live = None
        self._t = 0
        self._slices = None

    # ---- checkpointing: the moments and the step count live outside torch's per-parameter `state` ----------
    def state_dict(self):
        sd = super().state_dict()
        t = int(self._state[0].item()) if (self.capturable and self._state is not None) else self._t
        sd["dgmk_flat"] = {"step": t, "exp_avg": None if self._m is None else self._m.detach().clone(),
                           "exp_avg_sq": None if self._v is None else self._v.detach().clone()}
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        flat = state_dict.pop("dgmk_flat", None)
        super().load_state_dict(state_dict)
        if flat is not None:
            dev = self.net.flat_theta().device
            self._t = int(flat["step"])
            self._m = None if flat["exp_avg"] is None else flat["exp_avg"].to(dev).clone()
            self._v = None if flat["exp_avg_sq"] is None else flat["exp_avg_sq"].to(dev).clone()
            if self.capturable:
                self._state = torch.zeros(2, dtype=torch.int64, device=dev)
                self._state[0] = self._t

    def _flat_grad(self):
        """The flat gradient: zero-copy when every .grad is a view of one [P] buffer
        (what the fused steps hand out), otherwise gathered."""
        slices = self._slices
        if slices is None or slices[0][0] is not self.net._plist[0]:
            slices = self._slices = self.net.param_slices()
        live = [(p, off, n) for p, off, n, _ in slices if p.grad is not None]
        if not live:
            return None, None
        P = self.net.flat_theta().numel()
        base = live[0][0].grad._base
        if base is not None and base.numel() == P and base.dtype == torch.float32 and base.is_contiguous() \
                and all(p.grad._base is base and p.grad.data_ptr() == base.data_ptr() + 4 * off
                        for p, off, _ in live):
            flat = base
        else:
            flat = torch.zeros(P, dtype=torch.float32, device=self.net.flat_theta().device)
            for p, off, n in live:
                flat[off:off + n] = p.grad.reshape(-1)
        # live mask (parameters whose grad is None are skipped, like torch.optim.Adam): rebuilt only when the
        # set of parameters with a gradient changes -- the B = 32..256 loops call this every ~100 us
        key = tuple(p.grad is not None for p, *_ in slices)
        if self._live is None or self._live[0] != key or self._live[1].device != flat.device:
            mask = torch.zeros(P, dtype=torch.uint8)
            for p, off, n, _ in slices:
                if p.grad is not None:
                    mask[off:off + n] = 1
            self._live = (key, mask.to(flat.device))
        return flat, self._live[1]

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        theta = self.net.flat_theta()
        flat, mask = self._flat_grad()
        if flat is None:
            return loss
        if self._m is None or self._m.device != theta.device:
            self._m, self._v = torch.zeros_like(theta), torch.zeros_like(theta)
        g = self.param_groups[0]
        if self.capturable:
            if self._state is None or self._state.device != theta.device:
                self._state = torch.zeros(2, dtype=torch.int64, device=theta.device)
            kernels.adam_step_dev(theta, self._m, self._v, flat, mask, g["lr"], g["betas"][0], g["betas"][1],
                                  g["eps"], self._state)
            return loss
        self._t += 1
        kernels.adam_step(theta, self._m, self._v, flat, mask, g["lr"], g["betas"][0], g["betas"][1],
                          g["eps"], self._t)
        return loss
