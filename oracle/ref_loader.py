"""Import the UNMODIFIED reference modules from oracle/_ref/ (see oracle/vendor_ref.py).

TEST / BENCH INFRASTRUCTURE ONLY: used by `bench.py --impl reference`, by bench.py's `cpu_baseline` and
`cuda_eager_baseline` legs and by tests.  matplotlib is not installed in this image; the reference only
needs it importable (heat.py:20-21,31; fredholm.py:35-37), so three stub modules are registered first
(the same stub oracle/make_golden.py uses).
"""
from __future__ import annotations

import os
import sys
import types

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
NAMES = ("neural_networks", "dgm_net", "heat", "simple_ode", "fitzhugh_nagumo", "fredholm")


def available():
    return all(os.path.exists(os.path.join(REF_DIR, n + ".py")) for n in NAMES)


def load():
    """-> namespace with .neural_networks, .dgm_net, .heat, .simple_ode, .fitzhugh_nagumo, .fredholm."""
    if not available():
        raise FileNotFoundError(f"{REF_DIR} is empty: run `python oracle/vendor_ref.py` where /root/reference exists")
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
            import matplotlib.pylab  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            pylab = types.ModuleType("matplotlib.pylab")
            pylab.rcParams = {}
            style = types.ModuleType("matplotlib.style")
            style.use = lambda *a, **k: None
            mpl.pylab, mpl.style = pylab, style
            sys.modules.update({"matplotlib": mpl, "matplotlib.pylab": pylab, "matplotlib.style": style})
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import importlib
    import contextlib
    import io
    ns = types.SimpleNamespace()
    with contextlib.redirect_stdout(io.StringIO()):
        for n in NAMES:
            setattr(ns, n, importlib.import_module(n))
    return ns
