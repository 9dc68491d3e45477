"""Recipe: place the reference's own hot-path modules under oracle/_ref/ (git-ignored, NOT gpurun-ignored).

TEST / BENCH INFRASTRUCTURE ONLY.  The reference (gdetor/differential_equations_dnn) is pure Python: there
is nothing to compile.  /root/reference does not exist on the GPU box, so -- exactly like a compiled
`oracle/_ref/*.so` would -- the unmodified files travel there as a build OUTPUT of this recipe, never as
part of the repository's history (`oracle/_ref/` is listed in .gitignore).  `__graft_entry__.build()` runs
this whenever /root/reference is present; `bench.py --impl reference` and the `cuda_eager_baseline` leg
then EXECUTE the real reference (`heat.dgm_loss_func`, `dgm_net.DGM`, ... through oracle/ref_loader.py)
instead of the port in oracle/ref_port.py.

    python oracle/vendor_ref.py            # copies, prints the manifest
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
# the files SURVEY 8(a) cites for the hot path (+ the two-line timer decorator they import)
FILES = ("neural_networks.py", "dgm_net.py", "heat.py", "simple_ode.py", "fitzhugh_nagumo.py", "fredholm.py",
         "auxiliary_funs.py")


def vendor(verbose=True):
    if not os.path.isdir(REF):
        return None
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for f in FILES:
        src = os.path.join(REF, f)
        shutil.copyfile(src, os.path.join(DST, f))
        manifest[f] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    json.dump({"source": REF, "sha256": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"vendored {len(FILES)} reference files into {DST}")
    return manifest


if __name__ == "__main__":
    sys.exit(0 if vendor() is not None else 1)
