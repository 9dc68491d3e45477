"""Which chunk of which tile did a wrong result come from?  A[r, k] = tile(r) + 1000 * chunk(k) + row_in_tile/1000, W[n, k] = [k == 32 (n % 12)]:
C[r, n] must be A[r, 32 (n % 12)]; any other value names its source."""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from differential_equations_dnn_b200 import _cabi
lib = _cabi.load()
lib.dgmk_gemm_tc_probe.restype = C.c_int
lib.dgmk_gemm_tc_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_void_p]
N, K, ld = 128, 384, 512
M = 74 * 128 * 4
r = torch.arange(M, device="cuda")
A = torch.zeros(M, ld, device="cuda")
kk = torch.arange(K, device="cuda")
PAT = int(sys.argv[1]) if len(sys.argv) > 1 else 1
if PAT == 1:   # tile, chunk, position in the chunk
    A[:, :K] = (r // 128).float()[:, None] + 512.0 * (kk // 32).float()[None, :] + (kk % 32).float()[None, :] / 32
else:          # row in the tile, chunk, position in the chunk
    A[:, :K] = (r % 128).float()[:, None] + 128.0 * (kk // 32).float()[None, :] + (kk % 32).float()[None, :] / 32
w = torch.zeros(N, K, device="cuda")
for n in range(N):
    w[n, 32 * (n % 12) + (3 * (n // 12) + n) % 32] = 1.0
Bt3 = torch.cat([w.reshape(-1), w.reshape(-1), torch.zeros_like(w).reshape(-1)]).contiguous()
ref = (A[:, :K].double() @ w.double().t()).float()
for rep in range(2):
    Cm = torch.full((M, ld), -1.0, device="cuda")
    lib.dgmk_set_gemm_engine(1)
    lib.dgmk_gemm_tc_probe(A.data_ptr(), Bt3.data_ptr(), Cm.data_ptr(), M, N, K, ld, None)
    torch.cuda.synchronize()
    out = Cm[:, :N]
    bad = out != ref
    print(f"rep {rep}: bad elements {int(bad.sum())} of {M * N}")
    if int(bad.sum()):
        idx = bad.nonzero()
        tiles = torch.unique(idx[:, 0] // 128)
        print("  bad tiles", tiles.tolist()[:30])
        t = int(tiles[0])
        sub_bad = bad[t * 128:(t + 1) * 128]
        print(f"  tile {t} (pair {t % 74}, index {t // 74}): bad rows {torch.unique(sub_bad.nonzero()[:, 0]).tolist()[:140]}")
        print(f"  bad cols {torch.unique(sub_bad.nonzero()[:, 1]).tolist()}")
        rr = int(sub_bad.nonzero()[0, 0])
        print(f"  row {rr}: got {out[t * 128 + rr, :24].tolist()}")
        print(f"          ref {ref[t * 128 + rr, :24].tolist()}")
        # histogram of (got - ref) over the whole problem
        d = (out - ref)[bad]
        vals, cnt = torch.unique(d, return_counts=True)
        print("  differences:", list(zip(vals.tolist()[:20], cnt.tolist()[:20])))
