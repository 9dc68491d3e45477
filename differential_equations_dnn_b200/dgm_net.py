"""Drop-in for the reference's `dgm_net.py` (Sirignano & Spiliopoulos DGM network).

Same class names, constructor signatures, parameter names/shapes/init order and
state_dict keys as /root/reference/dgm_net.py:20-119; the arithmetic of `forward`
(dgm_net.py:53-68, :103-119) and of everything autograd derives from it runs in the
sm_100a kernels behind include/dgmk.h instead of torch ops.
"""
from torch import nn

from . import _cabi
from ._flat import FlatParamModule


class DGMLayer(nn.Module):
    """Parameter container of one LSTM-like layer (dgm_net.py:38-48).

    Attribute names follow the reference, including its crossed suffixes: the Z gate
    owns `Z_wg`/`Z_ug`, the G gate `G_wz`/`G_uz`.  It has no forward of
    its own: the owning DGM evaluates all layers in the fused pipeline.
    """

    def __init__(self, input_dim=1, hidden_size=50):
        super().__init__()
        for gate, (w, u) in (("Z", ("wg", "ug")), ("G", ("wz", "uz")), ("R", ("wr", "ur")), ("H", ("wh", "uh"))):
            setattr(self, f"{gate}_{w}", nn.Linear(hidden_size, hidden_size))
            setattr(self, f"{gate}_{u}", nn.Linear(input_dim, hidden_size, bias=False))

    def forward(self, x, s_old):
        raise RuntimeError("DGMLayer is evaluated by its owning DGM (fused kernels); call the DGM")


class DGM(FlatParamModule):
    """dgm_net.DGM(input_dim, output_dim, hidden_size, num_layers) -- dgm_net.py:75-101."""

    def __init__(self, input_dim=1, output_dim=1, hidden_size=1, num_layers=1):
        super().__init__()
        self.S_in = nn.Linear(input_dim, hidden_size)
        self.layers = nn.ModuleList([DGMLayer(input_dim, hidden_size) for _ in range(num_layers)])
        self.S_out = nn.Linear(hidden_size, output_dim)
        self._finish_init(_cabi.KIND_DGM_LINEAR, input_dim, output_dim, hidden_size, num_layers,
                          _cabi.ACT_TANH)
