"""Diagnostic: per-tensor error of the CUDA step on a golden fixture, for every GEMM engine, against the executed
reference in FP32 and (where the fixture keeps it) FP64.   python tools/diag_golden.py heat heat_dgm_h128l3_b203"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from conftest import golden, rel  # noqa: E402
from differential_equations_dnn_b200 import _cabi, kernels as K  # noqa: E402
from test_gpu_kernels import run, desc_of  # noqa: E402

prob, name = sys.argv[1], sys.argv[2]
g = golden(name)
lib = _cabi.load()
outs = {}
for eng in (1, 2, 0):
    lib.dgmk_set_gemm_engine(eng)
    outs[eng] = run(K, prob, g)
lib.dgmk_set_gemm_engine(1)
has64 = "grad_f64" in g
print("loss ref32 %.9g ref64 %.9g" % (float(g["loss"]), float(g["loss_f64"])), {e: "%.9g" % outs[e][0] for e in outs})
for i, (off, r, c, live) in enumerate(K.param_layout(desc_of(g))):
    n = r * max(c, 1)
    s = slice(off, off + n)
    ref = g["grad"][s]
    if not live or np.linalg.norm(ref) == 0:
        continue
    row = "t%-3d n=%-6d |g| %.3e " % (i, n, np.linalg.norm(ref))
    for e in outs:
        row += " eng%d-vs-32 %.2e" % (e, rel(outs[e][1][s], ref))
        if has64:
            row += " -vs-64 %.2e" % rel(outs[e][1][s], g["grad_f64"][s])
    if has64:
        row += "  ref32-vs-64 %.2e" % rel(ref, g["grad_f64"][s])
    print(row)
