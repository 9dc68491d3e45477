// Jet (Taylor-mode) algebra shared by every element-wise stage of the pipeline.
//
// A jet is an array of C floats: [v, d_0..d_{ND-1}, p_0..p_{NP-1}] -- value,
// first derivatives along input coordinates 0..ND-1, and the second derivatives
// listed in the channel set.  These rules replace what the reference obtains
// with nested torch.autograd.grad(create_graph=True) (heat.py:73-85,
// simple_ode.py:54-58, fitzhugh_nagumo.py:74-84); SURVEY 7.1 states them and
// oracle/jets_np.py is the FP64 restatement they are tested against.
//
// Everything here is __host__ __device__ so that the same functors can be
// compiled by g++ into the test-only emulation harness (tests/host_emul).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DGMK_HD __host__ __device__ __forceinline__
#else
#define DGMK_HD inline
#endif

// Inside the resident-tile kernels (dgmk_tile.cu defines DGMK_TILE_TU) every activation / stash buffer a functor
// touches lives in shared memory: telling the compiler so turns its generic LD / ST (64-bit addresses) into LDS / STS
// with 32-bit address arithmetic.  A no-op everywhere else (host code, the layer-wise kernels of dgmk_cuda.cu).
#if defined(DGMK_TILE_TU) && defined(__CUDA_ARCH__)
#define DGMK_SMEM(p) __builtin_assume(__isShared(p))
#else
#define DGMK_SMEM(p) ((void)0)
#endif

namespace dgmk {

enum { ACT_RELU = 0, ACT_SIGMOID = 1, ACT_TANH = 2, ACT_LEAKY = 3 };
enum { KIND_MLP = 0, KIND_DGM_LINEAR = 1, KIND_DGM_RAW = 2 };
// channel-set ids (C ABI: jet_order/dimension select one of these)
enum { CS_V = 0, CS_D1O1 = 1, CS_HEAT = 2, CS_D2O1 = 3, CS_D1O2 = 4, CS_D2O2 = 5, CS_COUNT = 6 };

template <int ND_, int NP_, int I0 = 0, int J0 = 0, int I1 = 0, int J1 = 0, int I2 = 0, int J2 = 0>
struct ChanSet {
  static constexpr int ND = ND_;
  static constexpr int NP = NP_;
  static constexpr int C = 1 + ND_ + NP_;
  static DGMK_HD constexpr int pi(int q) { return q == 0 ? I0 : (q == 1 ? I1 : I2); }
  static DGMK_HD constexpr int pj(int q) { return q == 0 ? J0 : (q == 1 ? J1 : J2); }
};
using CsV = ChanSet<0, 0>;                          // value only
using CsD1O1 = ChanSet<1, 0>;                       // v, t            (ODE, FHN)
using CsHeat = ChanSet<2, 1, 0, 0>;                 // v, x, t, xx     (heat)
using CsD2O1 = ChanSet<2, 0>;                       // v, x, t
using CsD1O2 = ChanSet<1, 1, 0, 0>;                 // v, t, tt
using CsD2O2 = ChanSet<2, 3, 0, 0, 0, 1, 1, 1>;     // v, x, t, xx, xt, tt

DGMK_HD int cs_channels(int cs) {
  return cs == CS_V ? 1 : cs == CS_D1O1 ? 2 : cs == CS_HEAT ? 4 : cs == CS_D2O1 ? 3 : cs == CS_D1O2 ? 3 : 6;
}
DGMK_HD int cs_ndirs(int cs) {
  return cs == CS_V ? 0 : (cs == CS_D1O1 || cs == CS_D1O2) ? 1 : 2;
}

// index / d with a 32-bit divide whenever the index fits (64-bit division is ~10x the
// instructions and dominated the value-only element-wise kernels)
DGMK_HD int64_t idiv(int64_t a, int32_t d) {
  return (((uint64_t)a >> 32) == 0) ? (int64_t)((uint32_t)a / (uint32_t)d) : a / d;
}

// tf32 split used by the tensor-core GEMM tiles: hi = x rounded to 10 explicit mantissa
// bits (nearest, ties away -- cvt.rna.tf32.f32), lo = x - hi (exact in FP32)
DGMK_HD float tf32_round(float x) {
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
#else
  union { float f; uint32_t u; } v; v.f = x;
  v.u = (v.u + 0x1000u) & 0xFFFFE000u;
  return v.f;
#endif
}

// ---- activations ------------------------------------------------------------
// tanhf/expf are the accurate libm-grade device routines (no --use_fast_math):
// MUFU.TANH's 2^-11 relative error would break the 1e-5 gradient parity
// (SURVEY 7.3 H5).
template <int ACT>
DGMK_HD float act_value(float a) {
  if (ACT == ACT_TANH) return tanhf(a);
  if (ACT == ACT_SIGMOID) return 1.0f / (1.0f + expf(-a));
  if (ACT == ACT_RELU) return a > 0.f ? a : 0.f;
  return a > 0.f ? a : 0.01f * a;
}
// sigma', sigma'', sigma''' through the OUTPUT y ("a-form" stash keeps y, not a).
// relu / leaky: y > 0  <=>  a > 0 (strict, like aten threshold_backward).
template <int ACT>
DGMK_HD void act_derivs(float y, float& d1, float& d2, float& d3) {
  if (ACT == ACT_TANH) {
    d1 = 1.0f - y * y;
    d2 = -2.0f * y * d1;
    d3 = d1 * (6.0f * y * y - 2.0f);
  } else if (ACT == ACT_SIGMOID) {
    d1 = y * (1.0f - y);
    d2 = d1 * (1.0f - 2.0f * y);
    d3 = d1 * (1.0f - 6.0f * d1);
  } else if (ACT == ACT_RELU) {
    d1 = y > 0.f ? 1.f : 0.f;
    d2 = 0.f;
    d3 = 0.f;
  } else {
    d1 = y > 0.f ? 1.f : 0.01f;
    d2 = 0.f;
    d3 = 0.f;
  }
}

// a[0] holds the pre-activation value on entry; on exit y is the output jet and
// a[0] is untouched (callers store y[0] in its place: the a-form).
template <class CS, int ACT>
DGMK_HD void act_fwd(const float* a, float* y) {
  y[0] = act_value<ACT>(a[0]);
  float d1, d2, d3;
  act_derivs<ACT>(y[0], d1, d2, d3);
#pragma unroll
  for (int k = 0; k < CS::ND; ++k) y[1 + k] = d1 * a[1 + k];
#pragma unroll
  for (int q = 0; q < CS::NP; ++q) {
    const int c = 1 + CS::ND + q;
    y[c] = d1 * a[c] + d2 * a[1 + CS::pi(q)] * a[1 + CS::pj(q)];
  }
}
// a-form (yv, a_1..a_{C-1}) -> output jet
template <class CS, int ACT>
DGMK_HD void aform_to_jet(const float* af, float* y) {
  float d1, d2, d3;
  act_derivs<ACT>(af[0], d1, d2, d3);
  y[0] = af[0];
#pragma unroll
  for (int k = 0; k < CS::ND; ++k) y[1 + k] = d1 * af[1 + k];
#pragma unroll
  for (int q = 0; q < CS::NP; ++q) {
    const int c = 1 + CS::ND + q;
    y[c] = d1 * af[c] + d2 * af[1 + CS::pi(q)] * af[1 + CS::pj(q)];
  }
}
// cotangent of the pre-activation jet; af = a-form (af[0] = output value)
template <class CS, int ACT>
DGMK_HD void act_adj(const float* ybar, const float* af, float* abar) {
  float d1, d2, d3;
  act_derivs<ACT>(af[0], d1, d2, d3);
  float v = d1 * ybar[0];
#pragma unroll
  for (int k = 0; k < CS::ND; ++k) {
    abar[1 + k] = d1 * ybar[1 + k];
    v += d2 * af[1 + k] * ybar[1 + k];
  }
#pragma unroll
  for (int q = 0; q < CS::NP; ++q) {
    const int c = 1 + CS::ND + q, i = 1 + CS::pi(q), j = 1 + CS::pj(q);
    abar[c] = d1 * ybar[c];
    abar[i] += d2 * af[j] * ybar[c];
    abar[j] += d2 * af[i] * ybar[c];
    v += (d2 * af[c] + d3 * af[i] * af[j]) * ybar[c];
  }
  abar[0] = v;
}

// ---- products ---------------------------------------------------------------
template <class CS>
DGMK_HD void prod_fwd(const float* p, const float* q, float* r) {
  r[0] = p[0] * q[0];
#pragma unroll
  for (int k = 0; k < CS::ND; ++k) r[1 + k] = p[1 + k] * q[0] + p[0] * q[1 + k];
#pragma unroll
  for (int n = 0; n < CS::NP; ++n) {
    const int c = 1 + CS::ND + n, i = 1 + CS::pi(n), j = 1 + CS::pj(n);
    r[c] = p[c] * q[0] + p[i] * q[j] + p[j] * q[i] + p[0] * q[c];
  }
}
// pbar (+)= cotangent w.r.t. p of r = p*q; ACC selects accumulate vs overwrite
template <class CS, bool ACC>
DGMK_HD void prod_adj(const float* rbar, const float* q, float* pbar) {
  float t[CS::C];
  t[0] = q[0] * rbar[0];
#pragma unroll
  for (int k = 0; k < CS::ND; ++k) {
    t[0] += q[1 + k] * rbar[1 + k];
    t[1 + k] = q[0] * rbar[1 + k];
  }
#pragma unroll
  for (int n = 0; n < CS::NP; ++n) {
    const int c = 1 + CS::ND + n, i = 1 + CS::pi(n), j = 1 + CS::pj(n);
    t[0] += q[c] * rbar[c];
    t[i] += q[j] * rbar[c];
    t[j] += q[i] * rbar[c];
    t[c] = q[0] * rbar[c];
  }
#pragma unroll
  for (int c = 0; c < CS::C; ++c) pbar[c] = ACC ? pbar[c] + t[c] : t[c];
}

}  // namespace dgmk
