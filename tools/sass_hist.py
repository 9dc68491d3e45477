"""SASS opcode histogram of the built library, per kernel: the evidence that the contraction kernels are
Blackwell-native (tcgen05 MMA = UTCHMMA / UTCQMMA..., tensor-memory loads / stores = LDTM / STTM, bulk-copy
engine = UBLKCP, tcgen05.commit = UTCBAR, packed FP32 FMA = FFMA2).

    python tools/sass_hist.py [libdgmk.so] > profiles/r02_sass_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "differential_equations_dnn_b200", "csrc", "libdgmk.so")
OPS = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "FFMA", "HMMA", "MUFU", "LDS", "STS", "LDG", "STG",
       "RED", "ATOM", "LDL", "STL")


def histogram(lib=LIB):
    """{demangled kernel name: Counter(opcode -> count)}"""
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    names = {}
    out, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            out[cur][op] += 1
    mangled = list(out)
    if mangled:
        dem = subprocess.run(["c++filt"], input="\n".join(mangled), capture_output=True, text=True).stdout.splitlines()
        names = dict(zip(mangled, dem))
    return {names.get(k, k): v for k, v in out.items()}


def short(name, n=150):
    name = re.sub(r"\bdgmk::", "", name)
    name = re.sub(r"^void ", "", name)
    return name if len(name) <= n else name[:n - 3] + "..."


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    h = histogram(lib)
    tot = collections.Counter()
    for c in h.values():
        tot.update(c)
    print(f"# SASS opcode histogram of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a); {len(h)} kernels")
    print("# totals: " + "  ".join(f"{op} {tot[op]}" for op in OPS if tot[op]))
    print("# LDL / STL = local-memory (spill) loads / stores")
    print()
    cols = [op for op in OPS if tot[op]]
    print("%-150s %6s " % ("kernel", "instr") + " ".join("%7s" % c for c in cols))
    for name in sorted(h, key=lambda k: (-h[k]["UTCHMMA"], -sum(h[k].values()))):
        c = h[name]
        print("%-150s %6d " % (short(name), sum(c.values())) + " ".join("%7d" % c[op] for op in cols))


if __name__ == "__main__":
    main()
