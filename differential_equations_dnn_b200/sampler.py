"""On-device collocation sampler: Philox4x32-10 in one launch per draw.

The reference drivers draw their collocation points with `torch.rand` / `rand_like` -- two launches plus four `cat`s per
heat step (heat.py:125-134), k `rand_like` + k multiplies + a stack per Fredholm step (fredholm.py:66-67,100).  At the
reference's batch sizes those launches ARE the step.  `PhiloxSampler` replaces them with ONE kernel of this library
(`dgmk_sample_uniform` / `dgmk_sample_heat`, include/dgmk.h): counter-based, so a draw is a pure function of
(seed, stream, step, element) and the only state is a step counter in device memory that a captured CUDA graph can
advance by itself.  Statistical, not bitwise, parity with torch's stream (the drivers keep `sampler="torch"` as the
default); the generator is pinned bit for bit against oracle/philox_np.py in the tests.
"""
import torch

from . import kernels as K
from . import parallel


class PhiloxSampler:
    """seed: defaults to torch.initial_seed() (so `torch.manual_seed` steers it like it steers the reference's draws).
    Ranks of a data-parallel job draw from disjoint streams (stream id offset = rank * 256).  `step` is the device
    counter the kernels read; `_loop.graphed_loop(counter=sampler.step)` advances it once per iteration, an eager
    loop passes `step_add=i` instead."""

    def __init__(self, device, seed=None):
        self.device = device
        self.seed = int(torch.initial_seed() if seed is None else seed)
        self.step = torch.zeros(1, dtype=torch.int64, device=device)
        self.stream0 = parallel.rank() * 256

    def uniform(self, out, lo, hi, stream_id=0, step_add=0):
        return K.sample_uniform(out, lo, hi, self.seed, self.stream0 + stream_id, self.step, step_add)

    def heat(self, X, X0, XBD1, XBD2, xmax, tmax, xbd2, step_add=0):
        if self.stream0:   # dgmk_sample_heat uses streams 0 and 1: fold the rank into the seed instead
            seed = (self.seed + 0x9E3779B97F4A7C15 * parallel.rank()) & (2 ** 64 - 1)
        else:
            seed = self.seed
        K.sample_heat(X, X0, XBD1, XBD2, xmax, tmax, xbd2, seed, self.step, step_add)
