"""B200-native collocation training step for Deep-Galerkin solvers.

Drop-in for the data-parallel hot path of gdetor/differential_equations_dnn: the module
names mirror the reference's flat files (`neural_networks`, `dgm_net`, `heat`,
`simple_ode`, `fitzhugh_nagumo`, `fredholm`) so `from neural_networks import MLP`
becomes `from differential_equations_dnn_b200.neural_networks import MLP`.
All arithmetic runs in hand-written sm_100a CUDA behind the C ABI in include/dgmk.h;
there is no CPU path.
"""
__version__ = "0.1.0"
