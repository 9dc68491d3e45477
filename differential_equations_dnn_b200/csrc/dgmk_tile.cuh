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
// :136-141, simple_ode.py:41-63 + :96-104, fitzhugh_nagumo.py:53-97 + :135-143, fredholm.py:47-74 + :104-109 (blocks of
// points whose k quadrature nodes are walked in sub-tiles: dgmk_steps.h fredholm_block).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "dgmk_steps.h"
#include "dgmk_tile_params.h"

namespace dgmk {
namespace tk {

// What Pipeline<> reads of a context; n / pl refer to the kernel's __grid_constant__ parameter block
struct TileCtx {
  const NetDims& n; const PackedLayout& pl;
  float* Wp; float* Gp; float* part; int64_t part_n; float* Lp;
  __device__ TileCtx(const NetDims& n_, const PackedLayout& pl_) : n(n_), pl(pl_), Wp(nullptr), Gp(nullptr), part(nullptr), part_n(0), Lp(nullptr) {}
};

// The CTA's dynamic shared memory (every extern __shared__ array of a kernel names the same block).  The stage
// primitives below receive generic pointers into it; rebasing them on this symbol tells the compiler their address
// space, so that plain C++ accesses compile to LDS / STS (.128 for float4) with 32-bit addresses and full freedom to
// schedule -- neither the generic LD / ST of an unknown pointer nor order-pinned inline asm.
extern __shared__ __align__(16) float g_tile_smem[];
template <class T>
__device__ __forceinline__ T* shp(T* p) {
  return reinterpret_cast<T*>(g_tile_smem + (reinterpret_cast<const float*>(p) - reinterpret_cast<const float*>(g_tile_smem)));
}

// ---- CTA-cooperative stage primitives (free functions, NOT inlined: one copy per translation unit keeps the
// kernels' instruction footprint and compile time down; every kernel instantiation calls the same code) ---------

// C[M,N] (+)= A[M,K] B[K,N]; A, C in shared memory, B = packed weights (shared memory: BSH, or L2).  TM x 4 register
// tiles on packed FFMA2, the tile index walks N fastest (a warp shares its A rows by broadcast).
template <int TM, bool BSH>
__device__ __noinline__ void gemm_nn_impl(const float* __restrict__ A_, int lda, const float* __restrict__ B_, int ldb, float* __restrict__ C_,
                                          int ldc, int M, int N, int K, bool acc, bool sync) {
  const float* A = shp(A_);
  float* C = shp(C_);
  const float* B = BSH ? shp(B_) : B_;
  const int ncg = N >> 2, ntiles = ((M + TM - 1) / TM) * ncg;
  for (int t = threadIdx.x; t < ntiles; t += NT) {
    const int rg = t / ncg, cg = t - rg * ncg;
    const int r0 = rg * TM, n0 = cg * 4;
    const float* ar[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) ar[i] = A + (r0 + i < M ? r0 + i : M - 1) * lda;   // clamp: rows >= M are computed, not stored
    float2 c2[TM][2];
#pragma unroll
    for (int i = 0; i < TM; ++i) { c2[i][0] = make_float2(0.f, 0.f); c2[i][1] = make_float2(0.f, 0.f); }
    const float* bp = B + n0;
#pragma unroll 2
    for (int k0 = 0; k0 < K; k0 += 4) {
      float4 b4[4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (BSH) b4[kk] = *reinterpret_cast<const float4*>(bp + (k0 + kk) * ldb);
        else b4[kk] = __ldg(reinterpret_cast<const float4*>(bp + (k0 + kk) * ldb));
      }
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const float4 a4 = *reinterpret_cast<const float4*>(ar[i] + k0);
        float2 a2;
        a2 = make_float2(a4.x, a4.x);
        c2[i][0] = __ffma2_rn(a2, make_float2(b4[0].x, b4[0].y), c2[i][0]); c2[i][1] = __ffma2_rn(a2, make_float2(b4[0].z, b4[0].w), c2[i][1]);
        a2 = make_float2(a4.y, a4.y);
        c2[i][0] = __ffma2_rn(a2, make_float2(b4[1].x, b4[1].y), c2[i][0]); c2[i][1] = __ffma2_rn(a2, make_float2(b4[1].z, b4[1].w), c2[i][1]);
        a2 = make_float2(a4.z, a4.z);
        c2[i][0] = __ffma2_rn(a2, make_float2(b4[2].x, b4[2].y), c2[i][0]); c2[i][1] = __ffma2_rn(a2, make_float2(b4[2].z, b4[2].w), c2[i][1]);
        a2 = make_float2(a4.w, a4.w);
        c2[i][0] = __ffma2_rn(a2, make_float2(b4[3].x, b4[3].y), c2[i][0]); c2[i][1] = __ffma2_rn(a2, make_float2(b4[3].z, b4[3].w), c2[i][1]);
      }
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      if (r0 + i < M) {
        float4* p = reinterpret_cast<float4*>(C + (r0 + i) * ldc + n0);
        float4 v = make_float4(c2[i][0].x, c2[i][0].y, c2[i][1].x, c2[i][1].y);
        if (acc) { const float4 o = *p; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
        *p = v;
      }
    }
  }
  if (sync) __syncthreads();
}

// column sums with up to four weights per row:
//   out[(ldo > 0 ? e * ldo : e * N) + n] += sum_r Wt[r][e] * Mat[r * ldm + n]   (e < 4; Wt == nullptr: plain sum, e = 0)
// thread = (column n, row group g of G = NT / N).  Narrow matrices (N = 1, 4, ...: loss sums, grad b_out) have many
// groups inside one warp: those are combined by shuffles first, so that the second stage adds at most NT / 32 partials
// per output (summing 256 partials in one thread cost more than the sums themselves).  Fixed order throughout.
struct ColsumJob {
  const float* Mat; int ldm, N; const float* Wt; int M; float* out; int ldo;
  int G, Gs, NE; float* scratch;   // filled in by colsum_plan
};
// groups of a job and the scratch floats its partials need (cap: what is left of the scratch)
__device__ __forceinline__ int colsum_plan(ColsumJob& j, float* scratch, int cap) {
  j.NE = j.Wt ? 4 : 1;
  const bool narrow = j.N < 32 && (32 % j.N) == 0;
  j.G = NT / j.N;
  if (!narrow && j.G * j.NE * j.N > cap) j.G = cap / (j.NE * j.N);
  j.Gs = narrow ? NT / 32 : j.G;
  j.scratch = scratch;
  return j.Gs * j.NE * j.N;
}
// first half: every thread's running sums -> one partial per (output, group) in the job's scratch
__device__ __forceinline__ void colsum_partial(const ColsumJob& j) {
  const float* Mat = shp(j.Mat);
  const float* Wt = j.Wt ? shp(j.Wt) : nullptr;
  float* scratch = shp(j.scratch);
  const int N = j.N, NE = j.NE, G = j.G, M = j.M, ldm = j.ldm;
  const int tid = threadIdx.x;
  const bool narrow = N < 32 && (32 % N) == 0;
  const int n = tid % N, g = tid / N;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (g < G) {
    if (Wt) {
#pragma unroll 4
      for (int r = g; r < M; r += G) {
        const float v = Mat[r * ldm + n];
        const float4 w = *reinterpret_cast<const float4*>(Wt + r * 4);
        a0 = fmaf(w.x, v, a0); a1 = fmaf(w.y, v, a1); a2 = fmaf(w.z, v, a2); a3 = fmaf(w.w, v, a3);
      }
    } else {
#pragma unroll 4
      for (int r = g; r < M; r += G) a0 += Mat[r * ldm + n];
    }
  }
  if (narrow) {   // the 32 / N groups of a warp -> one partial per warp (whole warps: NT is a multiple of 32)
    for (int off = N; off < 32; off <<= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, off);
      if (Wt) {
        a1 += __shfl_xor_sync(0xffffffffu, a1, off);
        a2 += __shfl_xor_sync(0xffffffffu, a2, off);
        a3 += __shfl_xor_sync(0xffffffffu, a3, off);
      }
    }
    if ((tid & 31) < N) {
      float* s = scratch + ((tid >> 5) * NE) * N + n;
      s[0] = a0;
      if (Wt) { s[N] = a1; s[2 * N] = a2; s[3 * N] = a3; }
    }
  } else if (g < G) {
    float* s = scratch + (g * NE) * N + n;
    s[0] = a0;
    if (Wt) { s[N] = a1; s[2 * N] = a2; s[3 * N] = a3; }
  }
}
// second half (after a barrier): the partials of an output in group order, added to the accumulator
__device__ __forceinline__ void colsum_finish(const ColsumJob& j, int tid0) {
  const float* scratch = shp(j.scratch);
  const int N = j.N, NEN = j.NE * j.N;
  for (int i = (int)threadIdx.x - tid0; i >= 0 && i < NEN; i += NT) {
    float t = 0.f;
    for (int q = 0; q < j.Gs; ++q) t += scratch[q * NEN + i];
    const int e = i / N, c = i - e * N;
    j.out[(j.ldo > 0 ? e * j.ldo : e * N) + c] += t;
  }
}
// (more columns than threads -- hidden size 128, three gates: one column block of NT per round; the loop lives in
// this non-inlined function: in the callers' wrappers it cost the kernels ~400 bytes of spills)
__device__ __noinline__ void wcolsum_impl(const float* __restrict__ Mat_, int ldm, int N, const float* __restrict__ Wt_, int M, float* out,
                                          int ldo, float* scratch_) {
  for (int n0 = 0; n0 < N; n0 += NT) {
    ColsumJob j; j.Mat = Mat_ + n0; j.ldm = ldm; j.N = N - n0 < NT ? N - n0 : NT; j.Wt = Wt_; j.M = M; j.out = out + n0;
    j.ldo = ldo > 0 ? ldo : N;
    colsum_plan(j, scratch_, SCRATCH_FLOATS);
    colsum_partial(j);
    __syncthreads();
    colsum_finish(j, 0);
    __syncthreads();
  }
}
// The three column sums at the head of a reverse pass in ONE stage (two barriers instead of six): the pass's loss rows
// (Lp == nullptr: none), grad W_out = UB^T S_L and grad b_out = UB^T E.  Same arithmetic and summation order as three
// wcolsum_impl calls; the second halves run on different warps (tid0 = 0 / 32 / 128 when the CTA has that many threads).
__device__ __noinline__ void wcolsum3_impl(const float* Lp, int loss_rows, float* loss_out, const float* SL, int Hp, const float* UB, int M,
                                           float* gw_out, const float* E, float* gb_out, float* scratch_) {
  ColsumJob jl, jw, jb;
  jl.Mat = Lp; jl.ldm = 1; jl.N = 1; jl.Wt = nullptr; jl.M = loss_rows; jl.out = loss_out; jl.ldo = 0;
  jw.Mat = SL; jw.ldm = Hp; jw.N = Hp; jw.Wt = UB; jw.M = M; jw.out = gw_out; jw.ldo = 0;
  jb.Mat = UB; jb.ldm = 4; jb.N = 4; jb.Wt = E; jb.M = M; jb.out = gb_out; jb.ldo = 0;
  int used = 0;
  if (Lp) used += colsum_plan(jl, scratch_, SCRATCH_FLOATS);
  used += colsum_plan(jb, scratch_ + used, SCRATCH_FLOATS - used);
  colsum_plan(jw, scratch_ + used, SCRATCH_FLOATS - used);
  if (Lp) colsum_partial(jl);
  colsum_partial(jb);
  colsum_partial(jw);
  __syncthreads();
  colsum_finish(jw, 0);                                  // 4 * Hp outputs: threads 0 .. 4 Hp - 1
  colsum_finish(jb, NT >= 256 ? NT - 32 : 0);            // 16 outputs
  if (Lp) colsum_finish(jl, NT >= 256 ? NT - 64 : 0);    // 1 output
  __syncthreads();
}

// out[N, Kd] += A^T S over the tile's M rows.  4 x 4 register tiles over (n, kd); when there are fewer tiles than
// threads the rows are split over thread groups that are combined through the scratch in group order.
// E != nullptr: outE[e * ldoE + n] += sum_m A[m, n] E[m, e] (grad[U | b] = Abar^T E) in the SAME stage -- its running
// sums go to the second scratch before the tiles start, its second half runs after the tiles' barrier.
__device__ __noinline__ void gemm_tn_impl(const float* __restrict__ A_, int lda, const float* __restrict__ S_, int lds, float* out, int N, int Kd,
                                          int M, float* scratch_, const float* E, float* outE, int ldoE) {
  const float* A = shp(A_);
  const float* S = shp(S_);
  float* scratch = shp(scratch_);
  ColsumJob je;
  if (SCRATCH2_FLOATS == 0) E = nullptr;   // (the caller runs the column sums as a stage of their own)
  if (E) {
    je.Mat = A_; je.ldm = lda; je.N = N; je.Wt = E; je.M = M; je.out = outE; je.ldo = ldoE;
    colsum_plan(je, scratch_ + SCRATCH_FLOATS, SCRATCH2_FLOATS);
    colsum_partial(je);
  }
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
      const float* ap = A + g * lda + n0;
      const float* sq = S + g * lds + k0;
      const int astep = G * lda, sstep = G * lds;
#pragma unroll 4
      for (int m = g; m < M; m += G) {
        const float4 a = *reinterpret_cast<const float4*>(ap);
        const float4 s = *reinterpret_cast<const float4*>(sq);
        ap += astep; sq += sstep;
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
        float4* p = reinterpret_cast<float4*>(out + (n0 + i) * Kd + k0);   // accumulators: shared memory or an L2 slot
        const float4 o = *p;
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        *p = v;
      }
    }
    if (G > 1) __syncthreads();   // (G > 1 implies a single pass of the t0 loop)
  }
  if (E) {
    if (G == 1) __syncthreads();   // the column-sum partials are visible (G > 1: the barriers above did that)
    colsum_finish(je, 0);
  }
  __syncthreads();
}

// u[r][m] = S[r, :] . W[m, :] (+ b[m] on value rows).  One thread per (row, output): the row and the weight row are
// read as float4 pieces starting at a lane-dependent column (a quarter warp then touches 8 different 16-byte columns
// of its 8 rows: conflict-free) with every load issued before the first FMA.
template <bool WSH>
__device__ __noinline__ void rowdot_impl(const float* __restrict__ S_, int lds, const float* __restrict__ W_, const float* __restrict__ b, float* U_, int M,
                                         int Hp, int o, int C) {
  const float* S = shp(S_);
  float* U = shp(U_);
  const float* W = WSH ? shp(W_) : W_;
  for (int t = threadIdx.x; t < M * 4; t += NT) {
    const int r = t >> 2, m = t & 3;
    float acc = 0.f;
    if (m < o) {
      const float* s = S + r * lds;
      const float* w = W + m * Hp;
      const int rot = ((t >> 2) & 7) * 4;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
      for (int q = 0; q < Hp; q += 4) {
        const int j = (rot + q) & (Hp - 1);   // Hp = 32 or 64
        const float4 sv = *reinterpret_cast<const float4*>(s + j);
        const float4 wv = WSH ? *reinterpret_cast<const float4*>(w + j) : __ldg(reinterpret_cast<const float4*>(w + j));
        a0 = fmaf(sv.x, wv.x, a0); a1 = fmaf(sv.y, wv.y, a1); a2 = fmaf(sv.z, wv.z, a2); a3 = fmaf(sv.w, wv.w, a3);
      }
      acc = (a0 + a1) + (a2 + a3);
      if ((r % C) == 0) acc += b[m];
    }
    U[r * 4 + m] = acc;
  }
  __syncthreads();
}

// The backend the orchestration templates see.  Every call ends with __syncthreads(): stages are separated exactly
// like kernel launches on a stream.  CSMASK / ACTMASK / MLP / DGM: the instantiation's live branches of the
// pipeline's run-time switches (BackendTraitsAll).
template <int CSMASK, int ACTMASK, bool MLP, bool DGM>
struct TileBackend {
  static constexpr bool cs_on(int id) { return (CSMASK >> id) & 1; }
  static constexpr bool act_on(int id) { return (ACTMASK >> id) & 1; }
  static constexpr bool mlp_on() { return MLP; }
  static constexpr bool dgm_on() { return DGM; }
  float* scratch;   // SCRATCH_TOTAL_FLOATS floats of shared memory (main scratch, then the A^T E scratch)
  int64_t hl_stride;
  // stage grouping (BackendTraitsAll::nosync): inside a group the element-wise / gemm_nn stages skip their barrier
  static constexpr bool kGroupsStages = true;
  bool grouped;
  __device__ __forceinline__ void nosync(bool on) { grouped = on; }
  bool w_shared, g_shared;   // packed weights / gradient accumulators live in shared memory (else L2)
  // stage timeline of CTA 0 (dgmk_tile_profile): (clock64, stage kind) after every stage; nullptr = off
  long long* prof; int prof_i, prof_n;
  __device__ __forceinline__ void stamp(int kind) {
    if (prof && threadIdx.x == 0 && prof_i < prof_n) { prof[2 * prof_i] = clock64(); prof[2 * prof_i + 1] = kind; ++prof_i; }
  }
  __device__ __forceinline__ bool inplace_rev() const { return true; }   // shared memory is the scarce resource
  // The coordinates of the tile's rows, copied once per pass from global memory into `coords` (every forward
  // stage that adds the input map re-reads them: from L2 / HBM that is a ~500-cycle load at the head of a stage
  // with one or two items per thread).  Rows keep their order; the three companion arrays become one block.
  float* coords; int coords_cap;
  __device__ __forceinline__ void stage_coords(XSrc& xs, int64_t rows_) {
    const int rows = (int)rows_, d = xs.d;
    if (rows * d > coords_cap) return;
    for (int i = threadIdx.x; i < rows * d; i += NT) {
      const int r = i / d, k = i - r * d;
      coords[i] = __ldg(xs.at(r) + k);
    }
    __syncthreads();
    // same block structure (the loss functors select targets by block), contiguous in the copy
    const int bd = (int)xs.block_rows * d;
    xs.p[0] = coords;
    if (xs.nptr > 1) { xs.p[1] = coords + bd; xs.p[2] = coords + 2 * bd; }
    else xs.block_stride = bd;
  }

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
  template <class CS, int ACT>
  __device__ void mlp_rev_fused(const float*, const float*, const float*, float*, int, int64_t) {}
  template <class CS, int ACT>
  __device__ void input_rev_fused(const float*, const float*, const dgmk::F4*, const float*, float*, int, int64_t) {}

  template <class F>
  __device__ __forceinline__ void ew(const F& f, int64_t n) {
    for (int i = threadIdx.x; i < (int)n; i += NT) f((int64_t)i);
    if (grouped) return;
    __syncthreads();
    stamp(1);
  }
  template <class F>
  __device__ __forceinline__ void ew4(const F& f, int64_t n) {
    const int n4 = (int)(n >> 2);
    for (int k = threadIdx.x; k < n4; k += NT) f.vec4((int64_t)k);
    if (grouped) return;
    __syncthreads();
    stamp(2);
  }
  // TM = 8 when 4-row tiles would need more than one round of the CTA's threads
  __device__ __forceinline__ void gemm_nn(const float* A, int64_t lda, const float* B, int64_t ldb, const float*, int64_t, float* C, int64_t ldc,
                                          int64_t M_, int N, int K, bool acc) {
    const int M = (int)M_;
    const bool big = ((M + 3) >> 2) * (N >> 2) > NT;
    const bool sync = !grouped;
    if (w_shared) {
      if (big) gemm_nn_impl<8, true>(A, (int)lda, B, (int)ldb, C, (int)ldc, M, N, K, acc, sync);
      else gemm_nn_impl<4, true>(A, (int)lda, B, (int)ldb, C, (int)ldc, M, N, K, acc, sync);
    } else {
      if (big) gemm_nn_impl<8, false>(A, (int)lda, B, (int)ldb, C, (int)ldc, M, N, K, acc, sync);
      else gemm_nn_impl<4, false>(A, (int)lda, B, (int)ldb, C, (int)ldc, M, N, K, acc, sync);
    }
    if (sync) stamp(3);
  }
  __device__ __forceinline__ void wcolsum_acc(const float* Mat, int64_t ldm, int N, const float* Wt, int64_t M, float* out, float*, int64_t,
                                              int64_t ldo = 0) {
    wcolsum_impl(Mat, (int)ldm, N, Wt, (int)M, out, (int)ldo, scratch);
    stamp(4);
  }
  __device__ __forceinline__ void wcolsum3(const float* Lp, int64_t loss_rows, float* loss_out, const float* SL, int Hp, const float* UB,
                                           int64_t M, float* gw_out, const float* E, float* gb_out) {
    wcolsum3_impl(Lp, (int)loss_rows, loss_out, SL, Hp, UB, (int)M, gw_out, E, gb_out, scratch);
    stamp(4);
  }
  // (+ outE[e * ldoE + n] += sum_m A[m, n] E[m, e]: grad[U | b] = Abar^T E)
  __device__ __forceinline__ void gemm_tn_acc(const float* A, int64_t lda, const float* S, int64_t lds, float* out, int N, int Kd, int64_t M,
                                              const float* E, float* outE, int64_t ldoE, float*, int64_t) {
    gemm_tn_impl(A, (int)lda, S, (int)lds, out, N, Kd, (int)M, scratch, E, outE, (int)ldoE);
    stamp(5);
    if (E && SCRATCH2_FLOATS == 0) { wcolsum_impl(A, (int)lda, N, E, (int)M, outE, (int)ldoE, scratch); stamp(6); }
  }
  __device__ __forceinline__ void rowdot(const float* S, int64_t lds, const float* W, const float* b, float* U, int64_t M, int Hp, int o, int C) {
    if (w_shared) rowdot_impl<true>(S, (int)lds, W, b, U, (int)M, Hp, o, C);
    else rowdot_impl<false>(S, (int)lds, W, b, U, (int)M, Hp, o, C);
    stamp(7);
  }
  __device__ __forceinline__ void zero(void* p, size_t bytes) {
    float* f = reinterpret_cast<float*>(p);
    for (int i = threadIdx.x; i < (int)(bytes >> 2); i += NT) f[i] = 0.f;
    __syncthreads();
  }
  __device__ void copy(void*, const void*, size_t) {}
};

// shared memory: [scratch][weights if w_smem][accumulators if g_smem][Lp][tile region]
template <int PROB, class BK>
// Register budget: the heat kernels (four-channel jets) want more than 128 registers and their tiles fill an SM's shared
// memory anyway; the value-only / first-order kernels fit 128, so two CTAs can share an SM when the tile is small enough.
__global__ void __launch_bounds__(NT, (PROB == PROB_HEAT || (DGMK_TILE_DGM_ONE_CTA && BK::dgm_on())) ? 1 : 2) tile_step_kernel(const __grid_constant__ TileParams prm) {
  float* sp = g_tile_smem;
  BK bk;
  bk.scratch = sp; sp += SCRATCH_TOTAL_FLOATS;
  bk.hl_stride = 0;
  bk.grouped = false;
  bk.w_shared = prm.w_smem != 0; bk.g_shared = prm.g_smem != 0;
  bk.prof = blockIdx.x == 0 ? prm.prof : nullptr; bk.prof_i = 0; bk.prof_n = prm.prof_n;
  bk.stamp(0);
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
  bk.coords = sp; bk.coords_cap = (int)prm.coord_floats; sp += prm.coord_floats;
  float* Ip = sp; sp += prm.ip_floats;   // Fredholm: running integral / residual gradient per point of the block
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
    else if (PROB == PROB_ODE) ode_like_chunk(pipe, cv, 0, prm.ode, p0, r);
    else fredholm_block(pipe, cv, 0, prm.fred, p0, r, prm.J, Ip);
    if (++in_seg == prm.flush_tiles && slot + 1 < prm.nslots_per_cta) {   // start a new FP32 accumulation segment
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
