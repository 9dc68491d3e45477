"""Diagnostic: resident-tile step (dgmk_set_tile_engine(1)) vs the layer-wise path (0) vs the golden reference."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from conftest import golden, golden_names, rel  # noqa: E402
from differential_equations_dnn_b200 import _cabi, kernels as K  # noqa: E402
from test_gpu_kernels import run, desc_of  # noqa: E402

lib = _cabi.load()
names = sys.argv[1:] or [n for p in ("heat_", "ode_", "fhn_", "fredholm_") for n in golden_names(p) if "driver" not in n]
for name in names:
    prob = name.split("_")[0]
    g = golden(name)
    H = int(g["spec"][3])
    if H > 64:
        continue
    res = {}
    for eng in (1, 0):
        lib.dgmk_set_tile_engine(eng)
        n0 = lib.dgmk_launch_count()
        res[eng] = run(K, prob, g)
        torch.cuda.synchronize()
        res[eng] += (lib.dgmk_launch_count() - n0,)
    lib.dgmk_set_tile_engine(1)
    worst = {e: 0.0 for e in res}
    for off, r, c, live in K.param_layout(desc_of(g)):
        n = r * max(c, 1)
        ref = g["grad"][off:off + n]
        if live and np.linalg.norm(ref) > 0 and n >= 8:
            for e in res:
                worst[e] = max(worst[e], rel(res[e][1][off:off + n], ref))
    print("%-28s loss ref %.7g tile %.7g layer %.7g | grad worst tile %.2e layer %.2e | launches tile %d layer %d" % (
        name, float(g["loss"]), res[1][0], res[0][0], worst[1], worst[0], res[1][2], res[0][2]), flush=True)
