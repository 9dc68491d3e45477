// Resident-tile training step for hidden sizes <= 64 (sm_100a): ONE persistent kernel per step.
//
// The layer-wise path (dgmk_pipeline.h driven from the host) launches ~150 kernels per step and round-trips every
// activation through HBM; at hidden size 32 that is 7-13 % of the FP32 roofline (profiles/r01_configs.txt) and
// ~170 us of launch latency at the reference's own batch sizes (B = 32 .. 256).  Here the chunk loop of the C ABI
// moves INTO the kernel: a chunk is a TILE of P collocation points whose whole activation stash (forward jets, a-form
// gates, reverse scratch) lives in shared memory, the packed weights are staged into shared memory once per CTA,
// and every stage of the SAME orchestration templates (Pipeline::forward / ::reverse, the *_chunk bodies of
// dgmk_steps.h, the functors of dgmk_ops.h) becomes a CTA-cooperative call separated by __syncthreads().  Weight
// gradients accumulate in per-CTA FP32 accumulators (shared memory when they fit, an L2-resident slot otherwise) and
// leave as one partial per CTA and segment of <= FLUSH tiles; a fixed-order FP64 second stage adds the partials
// (deterministic: tile -> CTA assignment is static).  Nothing but the point coordinates (40 B per heat row) is
// read from HBM and nothing but the partials is written.
//
// North star: "a tile of collocation points through every layer, weights held in shared memory ... a matching
// fused reverse pass accumulates parameter gradients with block reductions".  Reference replaced: heat.py:50-95 +
// :136-141, simple_ode.py:41-63 + :96-104, fitzhugh_nagumo.py:53-97 + :135-143.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "dgmk_steps.h"

namespace dgmk {
namespace tk {

constexpr int NT = 512;                 // threads per CTA (one CTA per SM)
constexpr int SCRATCH_FLOATS = 4096;    // cross-group reduction scratch (16 KB)
constexpr int SMEM_MAX = 232448;        // 227 KB opt-in limit per CTA on sm_100
constexpr int FLUSH_TILES = 256;        // tiles per FP32 accumulation segment

// What Pipeline<> reads of a context; n / pl refer to the kernel's __grid_constant__ parameter block
struct TileCtx {
  const NetDims& n; const PackedLayout& pl;
  float* Wp; float* Gp; float* part; int64_t part_n; float* Lp;
  __device__ TileCtx(const NetDims& n_, const PackedLayout& pl_) : n(n_), pl(pl_), Wp(nullptr), Gp(nullptr), part(nullptr), part_n(0), Lp(nullptr) {}
};

// CTA-cooperative implementations of the backend primitives.  Every call ends with __syncthreads(): stages are
// separated exactly like kernel launches on a stream.  CSMASK / ACTMASK / MLP / DGM: the instantiation's live
// branches of the pipeline's run-time switches (BackendTraitsAll).
template <int CSMASK, int ACTMASK, bool MLP, bool DGM>
struct TileBackend {
  static constexpr bool cs_on(int id) { return (CSMASK >> id) & 1; }
  static constexpr bool act_on(int id) { return (ACTMASK >> id) & 1; }
  static constexpr bool mlp_on() { return MLP; }
  static constexpr bool dgm_on() { return DGM; }
  float* scratch;   // SCRATCH_FLOATS floats of shared memory
  int64_t hl_stride;

  __device__ __forceinline__ void note_bytes(double) {}
  __device__ __forceinline__ bool lane_ok(int, int) const { return false; }
  // the tcgen05 units-on-lanes kernels belong to the hidden-size-128 path: never reached from here
  template <class CS, int ACT>
  __device__ void dgm_fwd_fused(const XSrc&, const float*, float*, const F4*, float*, float*, const float*, int, int64_t) {}
  template <class CS, int ACT>
  __device__ void dgm_rev2_fused(const float*, const float*, float*, float*, const float*, int, int64_t) {}
  template <class CS, class F>
  __device__ void dgm_rev1_e(const F&, const XSrc&, int64_t, float*, float*, int64_t) {}
  template <class CS, int ACT>
  __device__ void mlp_fwd_fused(const float*, float*, const F4*, float*, const float*, int, int64_t) {}
  __device__ void lane_store(const float*, int64_t, const float*, float*, int64_t, int, int64_t) {}

  template <class F>
  __device__ __forceinline__ void ew(const F& f, int64_t n) {
    for (int i = threadIdx.x; i < (int)n; i += NT) f((int64_t)i);
    __syncthreads();
  }
  template <class F>
  __device__ __forceinline__ void ew4(const F& f, int64_t n) {
    const int n4 = (int)(n >> 2);
    for (int k = threadIdx.x; k < n4; k += NT) f.vec4((int64_t)k);
    __syncthreads();
  }

  // C[M,N] (+)= A[M,K] B[K,N]; A, C in shared memory, B = packed weights (shared memory or L2).  4 x 4 register
  // tiles on packed FFMA2, the tile index walks N fastest (a warp shares its A rows by broadcast).
  __device__ void gemm_nn(const float* __restrict__ A, int64_t lda_, const float* __restrict__ B, int64_t ldb_, const float*, int64_t,
                          float* __restrict__ C, int64_t ldc_, int64_t M_, int N, int K, bool acc) {
    const int lda = (int)lda_, ldb = (int)ldb_, ldc = (int)ldc_, M = (int)M_;
    const int ncg = N >> 2, ntiles = ((M + 3) >> 2) * ncg;
    for (int t = threadIdx.x; t < ntiles; t += NT) {
      const int rg = t / ncg, cg = t - rg * ncg;
      const int r0 = rg * 4, n0 = cg * 4;
      const float* ar[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) ar[i] = A + (r0 + i < M ? r0 + i : M - 1) * lda;   // clamp: rows >= M are computed, not stored
      float2 c2[4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) { c2[i][0] = make_float2(0.f, 0.f); c2[i][1] = make_float2(0.f, 0.f); }
      const float* bp = B + n0;
      for (int k0 = 0; k0 < K; k0 += 4) {
        float4 a4[4], b4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a4[i] = *reinterpret_cast<const float4*>(ar[i] + k0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) b4[kk] = *reinterpret_cast<const float4*>(bp + (k0 + kk) * ldb);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float2 blo = make_float2(b4[kk].x, b4[kk].y), bhi = make_float2(b4[kk].z, b4[kk].w);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float av = kk == 0 ? a4[i].x : (kk == 1 ? a4[i].y : (kk == 2 ? a4[i].z : a4[i].w));
            const float2 a2 = make_float2(av, av);
            c2[i][0] = __ffma2_rn(a2, blo, c2[i][0]);
            c2[i][1] = __ffma2_rn(a2, bhi, c2[i][1]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (r0 + i < M) {
          float4* p = reinterpret_cast<float4*>(C + (r0 + i) * ldc + n0);
          float4 v = make_float4(c2[i][0].x, c2[i][0].y, c2[i][1].x, c2[i][1].y);
          if (acc) { const float4 o = *p; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
          *p = v;
        }
      }
    }
    __syncthreads();
  }

  // out[(i / cols) ...] column sums with up to four weights per row:
  //   out[(ldo > 0 ? e * ldo : e * N) + n] += sum_r Wt[r][e] * Mat[r * ldm + n]   (e < 4; Wt == nullptr: plain sum, e = 0)
  // thread = (column n, row group g of G = NT / N); groups are combined through the scratch in group order.
  __device__ void wcolsum_acc(const float* __restrict__ Mat, int64_t ldm_, int N, const float* __restrict__ Wt, int64_t M_, float* out,
                              float*, int64_t, int64_t ldo = 0) {
    const int ldm = (int)ldm_, M = (int)M_;
    const int NE = Wt ? 4 : 1;
    int G = NT / N;
    if (G * NE * N > SCRATCH_FLOATS) G = SCRATCH_FLOATS / (NE * N);
    if (G > M) G = M > 0 ? M : 1;
    const int tid = threadIdx.x, n = tid % N, g = tid / N;
    if (g < G) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int r = g; r < M; r += G) {
        const float v = Mat[r * ldm + n];
        if (Wt) {
          const float4 w = *reinterpret_cast<const float4*>(Wt + r * 4);
          a0 = fmaf(w.x, v, a0); a1 = fmaf(w.y, v, a1); a2 = fmaf(w.z, v, a2); a3 = fmaf(w.w, v, a3);
        } else {
          a0 += v;
        }
      }
      float* s = scratch + (g * NE) * N + n;
      s[0] = a0;
      if (Wt) { s[N] = a1; s[2 * N] = a2; s[3 * N] = a3; }
    }
    __syncthreads();
    for (int i = tid; i < NE * N; i += NT) {
      float t = 0.f;
      for (int q = 0; q < G; ++q) t += scratch[q * NE * N + i];
      const int e = i / N, c = i - e * N;
      out[(ldo > 0 ? (int64_t)e * ldo : (int64_t)e * N) + c] += t;
    }
    __syncthreads();
  }

  // out[N, Kd] += A^T S over the tile's M rows (+ outE[e * ldoE + n] += sum_m A[m, n] E[m, e]).  4 x 4 register tiles
  // over (n, kd); when there are fewer tiles than threads the rows are split over thread groups that are combined
  // through the scratch in group order.
  __device__ void gemm_tn_acc(const float* __restrict__ A, int64_t lda_, const float* __restrict__ S, int64_t lds_, float* out, int N, int Kd,
                              int64_t M_, const float* __restrict__ E, float* outE, int64_t ldoE, float*, int64_t) {
    const int lda = (int)lda_, lds = (int)lds_, M = (int)M_;
    const int nkg = Kd >> 2, ntiles = (N >> 2) * nkg;
    int G = NT / ntiles;
    if (G < 1) G = 1;
    while (G > 1 && (G - 1) * N * Kd > SCRATCH_FLOATS) --G;
    for (int t0 = 0; t0 < ntiles * G; t0 += NT) {
      const int t = t0 + threadIdx.x;
      const int g = t / ntiles, tt = t - g * ntiles;
      const bool live = t < ntiles * G;
      const int ng = tt / nkg, kg = tt - ng * nkg;
      const int n0 = ng * 4, k0 = kg * 4;
      float2 c2[4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) { c2[i][0] = make_float2(0.f, 0.f); c2[i][1] = make_float2(0.f, 0.f); }
      if (live) {
        for (int m = g; m < M; m += G) {
          const float4 a = *reinterpret_cast<const float4*>(A + m * lda + n0);
          const float4 s = *reinterpret_cast<const float4*>(S + m * lds + k0);
          const float2 slo = make_float2(s.x, s.y), shi = make_float2(s.z, s.w);
          float2 a2;
          a2 = make_float2(a.x, a.x); c2[0][0] = __ffma2_rn(a2, slo, c2[0][0]); c2[0][1] = __ffma2_rn(a2, shi, c2[0][1]);
          a2 = make_float2(a.y, a.y); c2[1][0] = __ffma2_rn(a2, slo, c2[1][0]); c2[1][1] = __ffma2_rn(a2, shi, c2[1][1]);
          a2 = make_float2(a.z, a.z); c2[2][0] = __ffma2_rn(a2, slo, c2[2][0]); c2[2][1] = __ffma2_rn(a2, shi, c2[2][1]);
          a2 = make_float2(a.w, a.w); c2[3][0] = __ffma2_rn(a2, slo, c2[3][0]); c2[3][1] = __ffma2_rn(a2, shi, c2[3][1]);
        }
        if (g > 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(scratch + (g - 1) * N * Kd + (n0 + i) * Kd + k0) = make_float4(c2[i][0].x, c2[i][0].y, c2[i][1].x, c2[i][1].y);
        }
      }
      if (G > 1) __syncthreads();
      if (live && g == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 v = make_float4(c2[i][0].x, c2[i][0].y, c2[i][1].x, c2[i][1].y);
          for (int q = 1; q < G; ++q) {
            const float4 o = *reinterpret_cast<const float4*>(scratch + (q - 1) * N * Kd + (n0 + i) * Kd + k0);
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
          }
          float4* p = reinterpret_cast<float4*>(out + (n0 + i) * Kd + k0);
          const float4 o = *p;
          v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
          *p = v;
        }
      }
      if (G > 1) __syncthreads();   // (G > 1 implies a single pass of the t0 loop)
    }
    __syncthreads();
    if (E) wcolsum_acc(A, lda_, N, E, M_, outE, nullptr, 0, ldoE);
  }

  // u[r][m] = S[r, :] . W[m, :] (+ b[m] on value rows); one thread per row, the column index rotated by the lane so
  // that a warp's 32 rows hit 32 different banks
  __device__ void rowdot(const float* __restrict__ S, int64_t lds_, const float* __restrict__ W, const float* __restrict__ b, float* U, int64_t M_, int Hp,
                         int o, int C) {
    const int lds = (int)lds_, M = (int)M_;
    for (int r = threadIdx.x; r < M; r += NT) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const float* s = S + r * lds;
      int j = threadIdx.x & 31;   // Hp is a multiple of 32
      for (int q = 0; q < Hp; ++q) {
        const float sv = s[j];
#pragma unroll
        for (int m = 0; m < 4; ++m)
          if (m < o) acc[m] = fmaf(sv, W[m * Hp + j], acc[m]);
        j = (j + 1 == Hp) ? 0 : j + 1;
      }
      const bool vrow = (r % C) == 0;
      float4 out;
      out.x = acc[0] + (vrow ? b[0] : 0.f);
      out.y = (1 < o) ? acc[1] + (vrow ? b[1] : 0.f) : 0.f;
      out.z = (2 < o) ? acc[2] + (vrow ? b[2] : 0.f) : 0.f;
      out.w = (3 < o) ? acc[3] + (vrow ? b[3] : 0.f) : 0.f;
      *reinterpret_cast<float4*>(U + r * 4) = out;
    }
    __syncthreads();
  }
  __device__ void zero(void*, size_t) {}
  __device__ void copy(void*, const void*, size_t) {}
};

enum { PROB_HEAT = 0, PROB_ODE = 1 };   // ODE covers simple_ode and FitzHugh-Nagumo (OdeArgs::fhn)

struct TileParams {
  NetDims n; PackedLayout pl;
  const float* Wp;        // packed weights (plain copy), global memory
  float* slots;           // [nslots][g_total] zero-initialised partial gradient accumulators
  int64_t B;              // points of this launch
  int32_t P;              // points per tile
  int32_t nslots_per_cta;
  int32_t w_smem, g_smem; // stage the weights / keep the accumulators in shared memory
  uint32_t w_floats, g_floats, lp_floats, tile_bytes;
  HeatArgs heat;
  OdeArgs ode;
};

// shared memory: [scratch][weights if w_smem][accumulators if g_smem][Lp][tile region]
template <int PROB, class BK>
__global__ void __launch_bounds__(NT, 1) tile_step_kernel(const __grid_constant__ TileParams prm) {
  extern __shared__ __align__(16) float smem[];
  float* sp = smem;
  BK bk;
  bk.scratch = sp; sp += SCRATCH_FLOATS;
  bk.hl_stride = 0;
  TileCtx c(prm.n, prm.pl);
  const int tid = threadIdx.x;
  if (prm.w_smem) {
    c.Wp = sp;
    const float4* src = reinterpret_cast<const float4*>(prm.Wp);
    float4* dst = reinterpret_cast<float4*>(sp);
    for (uint32_t i = tid; i < prm.w_floats / 4; i += NT) dst[i] = __ldg(src + i);
    sp += prm.w_floats;
  } else {
    c.Wp = const_cast<float*>(prm.Wp);
  }
  float* gsm = nullptr;
  int slot = 0;
  float* slot0 = prm.slots + (size_t)blockIdx.x * prm.nslots_per_cta * prm.g_floats;
  if (prm.g_smem) {
    gsm = sp; sp += prm.g_floats;
    for (uint32_t i = tid; i < prm.g_floats; i += NT) gsm[i] = 0.f;
    c.Gp = gsm;
  } else {
    c.Gp = slot0;
  }
  c.Lp = sp; sp += prm.lp_floats;
  char* region = reinterpret_cast<char*>(sp);
  __syncthreads();

  Pipeline<BK, TileCtx> pipe(bk, c);
  const int64_t ntiles = (prm.B + prm.P - 1) / prm.P;
  int in_seg = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t p0 = tile * prm.P;
    const int64_t r = (prm.B - p0 < prm.P) ? prm.B - p0 : prm.P;
    Carver cv(region, prm.tile_bytes);
    if (PROB == PROB_HEAT) heat_chunk(pipe, cv, 0, prm.heat, p0, r);
    else ode_like_chunk(pipe, cv, 0, prm.ode, p0, r);
    if (++in_seg == FLUSH_TILES && slot + 1 < prm.nslots_per_cta) {   // start a new FP32 accumulation segment
      if (prm.g_smem) {
        float* dst = slot0 + (size_t)slot * prm.g_floats;
        for (uint32_t i = tid; i < prm.g_floats; i += NT) { dst[i] = gsm[i]; gsm[i] = 0.f; }
        __syncthreads();
      } else {
        c.Gp = slot0 + (size_t)(slot + 1) * prm.g_floats;
      }
      ++slot; in_seg = 0;
    }
  }
  if (prm.g_smem) {
    float* dst = slot0 + (size_t)slot * prm.g_floats;
    for (uint32_t i = tid; i < prm.g_floats; i += NT) dst[i] = gsm[i];
  }
}

}  // namespace tk
}  // namespace dgmk
