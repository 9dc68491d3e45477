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
        owner = getattr(params[0], "_dgmk_flat", (None, 0))[0]
        self.net = owner() if owner is not None else None
        if self.net is None or len(params) != len(self.net.param_slices()):
            raise ValueError("FusedAdam needs net.parameters() of a differential_equations_dnn_b200 network")
        self._m = self._v = self._live = None
        self._t = 0

    def _flat_grad(self):
        """The flat gradient: zero-copy when every .grad is a view of one [P] buffer
        (what the fused steps hand out), otherwise gathered."""
        slices = self.net.param_slices()
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
        mask = torch.zeros(P, dtype=torch.uint8)
        for p, off, n, _ in slices:
            if p.grad is not None:
                mask[off:off + n] = 1
        return flat, mask

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        theta = self.net.flat_theta()
        flat, mask = self._flat_grad()
        if flat is None:
            return loss
        if self._m is None or self._m.device != theta.device:
            self._m, self._v = torch.zeros_like(theta), torch.zeros_like(theta)
        key = mask.numpy().tobytes()
        if self._live is None or self._live[0] != key:
            self._live = (key, mask.to(theta.device))
        g = self.param_groups[0]
        if self.capturable:
            if self._state is None or self._state.device != theta.device:
                self._state = torch.zeros(2, dtype=torch.int64, device=theta.device)
            kernels.adam_step_dev(theta, self._m, self._v, flat, self._live[1], g["lr"], g["betas"][0], g["betas"][1],
                                  g["eps"], self._state)
            return loss
        self._t += 1
        kernels.adam_step(theta, self._m, self._v, flat, self._live[1], g["lr"], g["betas"][0], g["betas"][1],
                          g["eps"], self._t)
        return loss
