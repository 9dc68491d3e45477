"""ctypes binding of the C ABI declared in include/dgmk.h.

The shared library is built in-tree (differential_equations_dnn_b200/csrc/libdgmk.so,
see `build()` below / __graft_entry__.build) from hand-written sm_100a CUDA.  There is
no CPU implementation: `load()` raises if the library is missing or is not the CUDA
build, and every entry point rejects host pointers.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.path.join(CSRC, "libdgmk.so")
BACKEND = b"cuda-sm100a"

KIND_MLP, KIND_DGM_LINEAR, KIND_DGM_RAW = 0, 1, 2
ACT_RELU, ACT_SIGMOID, ACT_TANH, ACT_LEAKY = 0, 1, 2, 3
ACT_IDS = {"relu": ACT_RELU, "sigmoid": ACT_SIGMOID, "tanh": ACT_TANH, "leaky_relu": ACT_LEAKY}
WS_HEAT, WS_ODE, WS_FHN, WS_FREDHOLM, WS_JET0, WS_JET1, WS_JET2 = range(7)


class NetDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("input_dim", C.c_int32), ("output_dim", C.c_int32),
                ("hidden_size", C.c_int32), ("num_layers", C.c_int32), ("activation", C.c_int32),
                ("reserved", C.c_int32 * 2)]


class DgmkError(RuntimeError):
    pass


_P = C.c_void_p
_SIGNATURES = {
    "dgmk_version": (C.c_int, []),
    "dgmk_backend": (C.c_char_p, []),
    "dgmk_last_error": (C.c_char_p, []),
    "dgmk_param_count": (C.c_int64, [C.POINTER(NetDesc)]),
    "dgmk_param_layout": (C.c_int, [C.POINTER(NetDesc), C.c_int32, C.POINTER(C.c_int64),
                                    C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "dgmk_workspace_bytes": (C.c_size_t, [C.POINTER(NetDesc), C.c_int32, C.c_int64, C.c_int32]),
    "dgmk_heat_step": (C.c_int, [C.POINTER(NetDesc), _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int64,
                                 C.c_float, _P, _P, _P, C.c_size_t, _P]),
    "dgmk_ode_step": (C.c_int, [C.POINTER(NetDesc), _P, _P, _P, _P, C.c_int64, C.c_int64, _P, _P, _P,
                                C.c_size_t, _P]),
    "dgmk_fhn_step": (C.c_int, [C.POINTER(NetDesc), _P, _P, _P, _P, C.c_int64, C.c_int64, _P, _P, _P,
                                C.c_size_t, _P]),
    "dgmk_fredholm_step": (C.c_int, [C.POINTER(NetDesc), _P, _P, _P, C.c_int64, C.c_int32, C.c_int64,
                                     _P, _P, _P, C.c_size_t, _P]),
    "dgmk_jet_forward": (C.c_int, [C.POINTER(NetDesc), _P, _P, C.c_int64, C.c_int32, _P, _P, _P, _P,
                                   C.c_size_t, _P]),
    "dgmk_jet_reverse": (C.c_int, [C.POINTER(NetDesc), _P, _P, C.c_int64, C.c_int32, _P, _P, _P, _P, _P,
                                   C.c_size_t, _P]),
    "dgmk_eval": (C.c_int, [C.POINTER(NetDesc), _P, _P, C.c_int64, _P, _P, C.c_size_t, _P]),
    "dgmk_adam": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double,
                            C.c_int64, _P]),
    "dgmk_adam_dev": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double,
                                _P, _P]),
    "dgmk_sample_uniform": (C.c_int, [_P, C.c_int64, C.c_float, C.c_float, C.c_ulonglong, C.c_uint32, _P, C.c_longlong, _P]),
    "dgmk_sample_heat": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_ulonglong, _P,
                                   C.c_longlong, _P]),
}
# diagnostics exported only by the CUDA library (bench.py)
_CUDA_ONLY = {
    "dgmk_launch_count": (C.c_ulonglong, []),
    "dgmk_ffma_probe": (C.c_int, [_P, _P, C.c_int, C.c_int, _P]),
    "dgmk_gemm_probe": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int, C.c_int, C.c_int64, _P]),
    "dgmk_gemm_tc_probe": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int, C.c_int, C.c_int64, _P]),
    "dgmk_set_gemm_engine": (None, [C.c_int]),
    "dgmk_set_tile_engine": (None, [C.c_int]),
    "dgmk_tile_profile": (None, [_P, C.c_int]),
    "dgmk_set_tile_flush": (None, [C.c_int]),
    "dgmk_profile": (None, [C.c_int]),
    "dgmk_profile_read": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_double),
                                    C.POINTER(C.c_double)]),
}
EXPORTS = tuple(_SIGNATURES) + tuple(_CUDA_ONLY)


def bind(lib):
    """Attach argtypes/restype for every symbol include/dgmk.h declares."""
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype, fn.argtypes = res, args
    return lib


_LIB = None


def load():
    """Load the CUDA library; fail loudly when it is missing (no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise DgmkError(
                f"{LIB_PATH} not found: build it with `python __graft_entry__.py build` "
                "(nvcc, sm_100a). This package has no CPU or PyTorch fallback.")
        lib = bind(C.CDLL(LIB_PATH))
        for name, (res, args) in _CUDA_ONLY.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.dgmk_backend() != BACKEND:
            raise DgmkError(f"{LIB_PATH} reports backend {lib.dgmk_backend()!r}, expected {BACKEND!r}")
        _LIB = lib
    return _LIB


def check(rc, lib=None):
    if rc != 0:
        lib = lib or load()
        raise DgmkError(f"dgmk error {rc}: {lib.dgmk_last_error().decode()}")


def make_desc(kind, d, o, H, L, act=ACT_TANH):
    return NetDesc(kind, d, o, H, L, act, (C.c_int32 * 2)(0, 0))


NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC"]
SOURCES = ("dgmk_cuda.cu", "dgmk_tile.cu")   # translation units of libdgmk.so, compiled in parallel


def build(force=False, verbose=False):
    """nvcc-compile csrc/*.cu for sm_100a and link them into csrc/libdgmk.so (in-tree)."""
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(CSRC, "..", "..", "include", "dgmk.h"))
    if not force and os.path.exists(LIB_PATH) and all(
            os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    procs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(CSRC, src[:-3] + ".o")
        cmd = ["nvcc", *NVCC_FLAGS, "-c", "-o", obj, os.path.join(CSRC, src)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, cwd=CSRC)))
        objs.append(obj)
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH, *objs], check=True, cwd=CSRC)
    return LIB_PATH
