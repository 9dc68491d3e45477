"""Drop-in for the hot-path classes of the reference's `neural_networks.py`.

`MLP` (neural_networks.py:180-270, batch_norm=False branch), `DGMLayer` / `DGM`
(neural_networks.py:44-177) keep the reference's names, signatures, parameter
names/shapes, init distributions and RNG draw order, so the same seed gives the same
initial weights.  `forward` runs in the sm_100a kernels (include/dgmk.h).

Out of scope: `batch_norm=True` (couples rows across the batch, so there
is no per-point jet) and the unused `ResNet*` classes -- both raise.
"""
import torch
from torch import nn
from torch.nn.init import xavier_uniform_

from . import _cabi
from ._flat import FlatParamModule


def selectActivationFunction(name, beta=1.5):
    """neural_networks.py:24-41 (unknown names fall back to ReLU, as there)."""
    table = {"relu": nn.ReLU, "sigmoid": nn.Sigmoid, "tanh": nn.Tanh, "leaky_relu": nn.LeakyReLU}
    if name not in table:
        print("Activation not found!")
        name = "relu"
    return table[name]()


class DGMLayer(nn.Module):
    """Parameter container of neural_networks.DGMLayer (:67-96): raw [in,out] matrices
    used as `x @ U + s @ W + b`, xavier(gain=relu) init, zero [1,H] biases."""

    def __init__(self, input_size=1, output_size=1, func="relu"):
        super().__init__()
        self.input_size, self.output_size = input_size, output_size
        gain = nn.init.calculate_gain("relu")
        for n in ("Uz", "Ug", "Ur", "Uh"):
            setattr(self, n, nn.Parameter(xavier_uniform_(torch.ones([input_size, output_size]), gain=gain)))
        for n in ("Wz", "Wg", "Wr", "Wh"):
            setattr(self, n, nn.Parameter(xavier_uniform_(torch.ones([output_size, output_size]), gain=gain)))
        for n in ("bz", "bg", "br", "bh"):
            setattr(self, n, nn.Parameter(torch.zeros([1, output_size])))

    def forward(self, x, s):
        raise RuntimeError("DGMLayer is evaluated by its owning DGM (fused kernels); call the DGM")


class DGM(FlatParamModule):
    """neural_networks.DGM(input_dim, output_dim, hidden_size, num_layers, func).

    Faithful to the reference's quirks: the inner layers are ReLU
    whatever `func` says (:146) -- `func` only picks the activation after `x_in` (relu, else tanh) -- and `dgm1` is registered but never evaluated (:145) --
    its 12 tensors never receive a gradient, so Adam never moves them.
    """

    _dead_prefixes = ("dgm1.",)

    def __init__(self, input_dim=1, output_dim=1, hidden_size=1, num_layers=1, func="relu"):
        super().__init__()
        self.x_in = nn.Linear(input_dim, hidden_size)
        self.dgm1 = DGMLayer(input_dim, hidden_size, func=func)
        self.layers = nn.ModuleList([DGMLayer(input_dim, hidden_size) for _ in range(num_layers)])
        self.x_out = nn.Linear(hidden_size, output_dim)
        xavier_uniform_(self.x_in.weight)
        xavier_uniform_(self.x_out.weight)
        # func selects the activation after x_in only: relu, anything else tanh (:153-156)
        self._finish_init(_cabi.KIND_DGM_RAW, input_dim, output_dim, hidden_size, num_layers,
                          _cabi.ACT_RELU if func == "relu" else _cabi.ACT_TANH)


class MLP(FlatParamModule):
    """neural_networks.MLP(input_dim, output_dim, hidden_size, num_layers, batch_norm,
    activation) -- :184-228, init :247-270."""

    def __init__(self, input_dim=2, output_dim=1, hidden_size=50, num_layers=1, batch_norm=False,
                 activation="relu"):
        super().__init__()
        if batch_norm:
            raise NotImplementedError("batch_norm=True is outside the fused hot path (SURVEY 2.1, N4)")
        if activation not in _cabi.ACT_IDS:
            print("Activation not found!")
            activation = "relu"
        self.activation = activation
        print("No batch normalization")
        self.bn = nn.Identity()
        self.fc_in = nn.Linear(input_dim, hidden_size)
        self.layers = nn.ModuleList([nn.Linear(hidden_size, hidden_size) for _ in range(num_layers)])
        self.fc_out = nn.Linear(hidden_size, output_dim)
        self.act = selectActivationFunction(activation)
        self.reset()
        self._finish_init(_cabi.KIND_MLP, input_dim, output_dim, hidden_size, num_layers,
                          _cabi.ACT_IDS[activation])

    def reset(self):
        """Xavier (sigmoid/tanh) or Kaiming (relu/leaky_relu) uniform, :247-270."""
        stack = [self.fc_in, *self.layers]
        if self.activation in ("relu", "leaky_relu"):
            for lin in stack + [self.fc_out]:
                nn.init.kaiming_uniform_(lin.weight, nonlinearity=self.activation)
        else:
            gain = nn.init.calculate_gain(self.activation)
            for lin in stack:
                nn.init.xavier_uniform_(lin.weight, gain=gain)
            nn.init.xavier_uniform_(self.fc_out.weight)


def _out_of_scope(name):
    def ctor(*a, **k):
        raise NotImplementedError(f"{name} is not on the collocation hot path (SURVEY 2.1)")
    return ctor


ResidualBlock, ResNetLayer, ResNet = (_out_of_scope(n) for n in ("ResidualBlock", "ResNetLayer", "ResNet"))
