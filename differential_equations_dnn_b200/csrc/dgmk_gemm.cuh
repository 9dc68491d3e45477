// FP32 CUDA-core GEMM tiles for the layer-wise jet pipeline (sm_100a).
//
// Replaces the aten::mm / aten::addmm launches of the reference's forward,
// double-backward and backward sweeps (dgm_net.py:63-67,
// neural_networks.py:115-123,241-245).  FP32 FFMA, no TF32: the parity bar is
// 1e-5 per tensor.
//
//   gemm_nn : C[M,N] (+)= A[M,K] * B[K,N]        forward / data-gradient
//   gemm_tn : P[s][N,Kd]  = sum_{m in split s} A[m,N]^T * S[m,Kd]   weight-gradient
//             (+ PE[s][4,N] = A^T E for the input-map / bias gradients, fused)
//
// All operands row-major.  K, N multiples of 32 (hidden sizes are padded to
// 32), M arbitrary.  128 x BN CTA tile, 256 threads, 8 x (BN/16) register
// micro-tile, BK = 16 k-slices double-buffered in shared memory with register-
// staged prefetch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dgmk {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 16;
constexpr int GEMM_NT = 256;
constexpr int GEMM_APAD = 4;

template <int BN>
struct GemmTile {
  static constexpr int TN = BN / 16;          // columns per thread
  static constexpr int NB4 = BN / 4;          // float4 per B row
  static constexpr int B_F4 = GEMM_BK * NB4;  // float4 per B slice
  static constexpr int B_PER_T = (B_F4 + GEMM_NT - 1) / GEMM_NT;
};

// ---- inner product on one k-slice ------------------------------------------
// Packed FP32 FMA (fma.rn.f32x2 -> SASS FFMA2, sm_100+): the accumulators are column
// pairs, b comes as natural float2 pairs out of the float4 shared loads and a[i] is the
// scalar-broadcast operand (ptxas folds make_float2(a,a) into the .F32 operand form, no
// MOV).  Same IEEE fma per lane as FFMA; half the issue slots, which leaves room for the
// LDS / address instructions next to a saturated FMA pipe (profiles/r01_notes.md).
template <int BN>
__device__ __forceinline__ void mma_slice(const float (*As)[GEMM_BM + GEMM_APAD],
                                          const float (*Bs)[BN], int tx, int ty,
                                          float2 (&acc)[8][GemmTile<BN>::TN / 2]) {
  constexpr int TN = GemmTile<BN>::TN;
#pragma unroll
  for (int kk = 0; kk < GEMM_BK; ++kk) {
    float a[8];
    float2 b[TN / 2];
    *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
    *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
    if constexpr (TN == 8) {
      *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      *reinterpret_cast<float4*>(&b[2]) = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
    } else if constexpr (TN == 4) {
      *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
    } else {
      b[0] = *reinterpret_cast<const float2*>(&Bs[kk][tx * 2]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 a2 = make_float2(a[i], a[i]);
#pragma unroll
      for (int j = 0; j < TN / 2; ++j) acc[i][j] = __ffma2_rn(a2, b[j], acc[i][j]);
    }
  }
}

template <int BN>
__device__ __forceinline__ int col_of(int tx, int j) {
  constexpr int TN = GemmTile<BN>::TN;
  if constexpr (TN == 8) return (j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4);
  if constexpr (TN == 4) return tx * 4 + j;
  return tx * 2 + j;
}
__device__ __forceinline__ int row_of(int ty, int i) {
  return (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
}

// ---- C[M,N] (+)= A[M,K] B[K,N] ------------------------------------------------
// grid = (N / BN, ceil(M / 128)); blockIdx.x walks N so that the CTAs that share
// an A row-block are co-scheduled and A is served from L2 after the first read.
template <int BN, bool ACCUM>
__global__ void __launch_bounds__(GEMM_NT, 2)
gemm_nn_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
               float* __restrict__ C, int64_t ldc, int64_t M, int K) {
  using T = GemmTile<BN>;
  __shared__ __align__(16) float As[2][GEMM_BK][GEMM_BM + GEMM_APAD];
  __shared__ __align__(16) float Bs[2][GEMM_BK][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * GEMM_BM;
  const int n0 = blockIdx.x * BN;

  float2 acc[8][T::TN / 2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < T::TN / 2; ++j) acc[i][j] = make_float2(0.f, 0.f);

  float4 ra[2], rb[T::B_PER_T];
  auto load_g = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      int idx = tid + q * GEMM_NT;
      int r = idx >> 2, k4 = (idx & 3) * 4;
      int64_t m = m0 + r;
      ra[q] = (m < M) ? __ldg(reinterpret_cast<const float4*>(A + m * lda + k0 + k4))
                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int q = 0; q < T::B_PER_T; ++q) {
      int idx = tid + q * GEMM_NT;
      if (idx < T::B_F4) {
        int k = idx / T::NB4, n4 = (idx % T::NB4) * 4;
        rb[q] = __ldg(reinterpret_cast<const float4*>(B + (int64_t)(k0 + k) * ldb + n0 + n4));
      }
    }
  };
  auto store_s = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      int idx = tid + q * GEMM_NT;
      int r = idx >> 2, k4 = (idx & 3) * 4;
      As[buf][k4 + 0][r] = ra[q].x;
      As[buf][k4 + 1][r] = ra[q].y;
      As[buf][k4 + 2][r] = ra[q].z;
      As[buf][k4 + 3][r] = ra[q].w;
    }
#pragma unroll
    for (int q = 0; q < T::B_PER_T; ++q) {
      int idx = tid + q * GEMM_NT;
      if (idx < T::B_F4) {
        int k = idx / T::NB4, n4 = (idx % T::NB4) * 4;
        *reinterpret_cast<float4*>(&Bs[buf][k][n4]) = rb[q];
      }
    }
  };

  const int nk = K / GEMM_BK;
  load_g(0);
  store_s(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_g((kt + 1) * GEMM_BK);
    mma_slice<BN>(As[buf], Bs[buf], tx, ty, acc);
    if (kt + 1 < nk) {
      store_s(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t m = m0 + row_of(ty, i);
    if (m >= M) continue;
    float* crow = C + m * ldc + n0;
    if constexpr (T::TN == 8) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float4* p = reinterpret_cast<float4*>(crow + h * 64 + tx * 4);
        float4 v = make_float4(acc[i][h * 2].x, acc[i][h * 2].y, acc[i][h * 2 + 1].x, acc[i][h * 2 + 1].y);
        if constexpr (ACCUM) { float4 o = *p; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
        *p = v;
      }
    } else if constexpr (T::TN == 4) {
      float4* p = reinterpret_cast<float4*>(crow + tx * 4);
      float4 v = make_float4(acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y);
      if constexpr (ACCUM) { float4 o = *p; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
      *p = v;
    } else {
      float2* p = reinterpret_cast<float2*>(crow + tx * 2);
      float2 v = acc[i][0];
      if constexpr (ACCUM) { float2 o = *p; v.x += o.x; v.y += o.y; }
      *p = v;
    }
  }
}

// ---- P[z][N,Kd] = sum_{m in [z*rows_per_split, ...)} A[m, 0:N]^T S[m, 0:Kd] ---------
// grid = (Kd / BN, ceil(N / 128), splits).  Partials are reduced in a fixed order
// by reduce_partials (deterministic, FP64 accumulate): SURVEY 7.3 H4.
// When E != nullptr the CTAs of the first column tile also accumulate
//   PE[z][e][n] = sum_m A[m, n] * E[m, e]   (e < 4)
// from the A slices they already stage in shared memory: that is grad[U | b] = Abar^T E
// (SURVEY 7.1 "input map"), which a separate pass would have to re-read Abar for.
template <int BN>
__global__ void __launch_bounds__(GEMM_NT, 2)
gemm_tn_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ S, int64_t lds,
               float* __restrict__ P, int N, int Kd, int64_t M, int64_t rows_per_split,
               const float* __restrict__ E, float* __restrict__ PE) {
  using T = GemmTile<BN>;
  __shared__ __align__(16) float As[2][GEMM_BK][GEMM_BM + GEMM_APAD];
  __shared__ __align__(16) float Bs[2][GEMM_BK][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.y * GEMM_BM;  // output row (column of A)
  const int j0 = blockIdx.x * BN;       // output col (column of S)
  const int64_t mb = (int64_t)blockIdx.z * rows_per_split;
  const int64_t me = (mb + rows_per_split < M) ? mb + rows_per_split : M;
  const bool do_e = (E != nullptr) && (blockIdx.x == 0);
  __shared__ __align__(16) float Es[2][GEMM_BK][4];
  float4 re = make_float4(0.f, 0.f, 0.f, 0.f);
  float eacc0 = 0.f, eacc1 = 0.f;
  const int ei = tid & 127, eh = (tid >> 7) * 2;

  float2 acc[8][T::TN / 2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < T::TN / 2; ++j) acc[i][j] = make_float2(0.f, 0.f);

  float4 ra[2], rb[T::B_PER_T];
  auto load_g = [&](int64_t mm0) {
    if (do_e && tid < GEMM_BK) {
      int64_t m = mm0 + tid;
      re = (m < me) ? __ldg(reinterpret_cast<const float4*>(E) + m) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {  // 16 rows x 128 cols = 512 float4
      int idx = tid + q * GEMM_NT;
      int r = idx >> 5, c4 = (idx & 31) * 4;
      int64_t m = mm0 + r;
      ra[q] = (m < me && i0 + c4 < N)
                  ? __ldg(reinterpret_cast<const float4*>(A + m * lda + i0 + c4))
                  : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int q = 0; q < T::B_PER_T; ++q) {
      int idx = tid + q * GEMM_NT;
      if (idx < T::B_F4) {
        int r = idx / T::NB4, n4 = (idx % T::NB4) * 4;
        int64_t m = mm0 + r;
        rb[q] = (m < me) ? __ldg(reinterpret_cast<const float4*>(S + m * lds + j0 + n4))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  auto store_s = [&](int buf) {
    if (do_e && tid < GEMM_BK) *reinterpret_cast<float4*>(&Es[buf][tid][0]) = re;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      int idx = tid + q * GEMM_NT;
      int r = idx >> 5, c4 = (idx & 31) * 4;
      *reinterpret_cast<float4*>(&As[buf][r][c4]) = ra[q];
    }
#pragma unroll
    for (int q = 0; q < T::B_PER_T; ++q) {
      int idx = tid + q * GEMM_NT;
      if (idx < T::B_F4) {
        int r = idx / T::NB4, n4 = (idx % T::NB4) * 4;
        *reinterpret_cast<float4*>(&Bs[buf][r][n4]) = rb[q];
      }
    }
  };

  const int64_t nk = (me - mb + GEMM_BK - 1) / GEMM_BK;
  if (nk > 0) {
    load_g(mb);
    store_s(0);
    __syncthreads();
    for (int64_t kt = 0; kt < nk; ++kt) {
      const int buf = (int)(kt & 1);
      if (kt + 1 < nk) load_g(mb + (kt + 1) * GEMM_BK);
      mma_slice<BN>(As[buf], Bs[buf], tx, ty, acc);
      if (do_e) {
#pragma unroll
        for (int mm = 0; mm < GEMM_BK; ++mm) {
          const float av = As[buf][mm][ei];
          eacc0 = fmaf(av, Es[buf][mm][eh], eacc0);
          eacc1 = fmaf(av, Es[buf][mm][eh + 1], eacc1);
        }
      }
      if (kt + 1 < nk) {
        store_s(buf ^ 1);
        __syncthreads();
      }
    }
  }
  if (do_e && i0 + ei < N) {
    float* pe = PE + (int64_t)blockIdx.z * 4 * N;
    pe[(int64_t)eh * N + i0 + ei] = eacc0;
    pe[(int64_t)(eh + 1) * N + i0 + ei] = eacc1;
  }
  float* Pz = P + (int64_t)blockIdx.z * N * Kd;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int r = i0 + row_of(ty, i);
    if (r >= N) continue;
#pragma unroll
    for (int j = 0; j < T::TN; ++j)
      Pz[(int64_t)r * Kd + j0 + col_of<BN>(tx, j)] = (j & 1) ? acc[i][j >> 1].y : acc[i][j >> 1].x;
  }
}

}  // namespace dgmk
