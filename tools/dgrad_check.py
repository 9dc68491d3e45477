"""K = 3H data-gradient GEMM stand-alone: the weight-resident kernel (dgmk_dgrad_res.cuh, engine 1) against the
round-1 streaming tile (engine 3) and an FP64 product -- accuracy at ragged sizes, then time at a bench-sized M."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from differential_equations_dnn_b200 import _cabi

lib = _cabi.load()
lib.dgmk_gemm_tc_probe.restype = C.c_int
lib.dgmk_gemm_tc_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_void_p]
N = 128


def tf32_hi(x):
    return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def weights(K, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    w = (torch.rand(N, K, device="cuda", generator=g) - 0.5) * 0.3
    hi = tf32_hi(w)
    return w, torch.cat([w.reshape(-1), hi.reshape(-1), (w - hi).reshape(-1)]).contiguous()


def run(engine, A, Bt3, M, K, ld):
    lib.dgmk_set_gemm_engine(engine)
    Cm = torch.full((M, ld), 7.0, device="cuda")
    rc = lib.dgmk_gemm_tc_probe(A.data_ptr(), Bt3.data_ptr(), Cm.data_ptr(), M, N, K, ld, None)
    torch.cuda.synchronize()
    assert rc == 0, lib.dgmk_last_error()
    assert bool((Cm[:, N:] == 7.0).all()), "wrote outside the result block"
    return Cm[:, :N]


ok = True
for (M, K) in ((128, 384), (1, 384), (257, 384), (1000, 384), (40000, 384), (5000, 256), (333, 160)):
    ld = 512
    g = torch.Generator(device="cuda").manual_seed(M)
    A = torch.randn(M, ld, device="cuda", generator=g)
    w, Bt3 = weights(K, 3)
    ref = A[:, :K].double() @ w.double().t()
    out = {}
    for e in (1, 3):
        c = run(e, A, Bt3, M, K, ld)
        out[e] = float((c.double() - ref).norm() / ref.norm())
    print(f"M={M} K={K}: resident {out[1]:.2e}  streaming {out[3]:.2e}")
    ok &= out[1] < 5e-7
M, K, ld = 1 << 21, 384, 512
A = torch.randn(M, ld, device="cuda")
w, Bt3 = weights(K, 3)
Cm = torch.zeros(M, 128, device="cuda")
for e in (3, 1):
    lib.dgmk_set_gemm_engine(e)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for it in range(6):
        if it == 1:
            ev[0].record()
        # lda = 512, ldc = 512 would need a [M, 512] result: the probe uses one ld, so time through a 512-wide alias
        rc = lib.dgmk_gemm_tc_probe(A.data_ptr(), Bt3.data_ptr(), A.data_ptr() + 4 * 384, M, N, K, ld, None)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 5
    gb = M * (K + N) * 4 / 1e9
    print(f"engine {e}: {ms:.3f} ms per {M} rows  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s  {gb / ms * 1e3:.0f} GB/s (A + C written)")
lib.dgmk_set_gemm_engine(1)
print("OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
