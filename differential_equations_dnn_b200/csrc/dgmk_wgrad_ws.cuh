// Warp-specialised tcgen05 weight-gradient kernel (sm_100a), 3xTF32 split:
//
//   P[s][n, k]     = sum_{m in segment s} A[m, n] * S[m, k]     n = output unit (128 per CTA), k = input unit
//   PE[s][h][e, n] = sum_m A[m, n] * E[m, e]                    input-map / bias gradients (CUDA cores;
//                                                               h = which half of every 32-row chunk)
// CTA z owns `nseg` consecutive segments of seg_rows rows (s = z*nseg + i): the pipeline runs
// straight through, but the FP32 partial sums are written out and restarted at every segment
// boundary so that no FP32 accumulation chain grows beyond seg_rows rows (the second stage adds
// the partials in FP64, fixed order).
//
// Same arithmetic as dgmk_gemm_tc_tn.cuh (chunks of 32 rows; lo*hi + hi*lo then hi*hi from zero in
// TMEM; chunk results summed in round-to-nearest registers; FP64 second stage over the splits),
// different machine mapping.  The streaming tile staged BOTH operands through shared memory with
// one buffer and one barrier per chunk -- load, split, store, sync, MMA and drain ran one after the
// other (~3400 cycles per 32-row chunk against 768 cycles of MMA).  Here:
//
//   * A^T is the M-side operand and lives in TENSOR MEMORY: UMMA row = output unit n = TMEM lane,
//     UMMA K = row of the chunk = TMEM column.  Loader thread n reads A[m0 .. m0+31, n] (a warp
//     instruction = 128 contiguous bytes of one row), splits into tf32 hi / lo and writes its own
//     lane with tcgen05.st: no transposition, no swizzle, no shared memory, and the MMA reads only
//     the S operand from shared memory (half the operand traffic of the SS form).  The same thread
//     accumulates A^T E for its unit, so that reduction needs no cross-warp step either.
//   * S is staged MN-major (SWIZZLE_128B_BASE32B, see dgmk_gemm_tc_tn.cuh) into a 3-deep ring.
//   * loaders, stagers, the MMA issuer and the drain warps run concurrently, decoupled by
//     mbarriers; two A buffers and two accumulator buffers in TMEM.
//
// 20 warps: 0-7 A loaders (thread = TMEM lane 32*(w%4)+lane, rows 16*(w/4).. of every chunk), 8-11 S
// stagers, 12-19 drain (warp w: lanes 32*(w%4).., columns 64*((w-12)/4)..); setmaxnreg moves
// registers from the loaders to the drain warps (64 accumulators each).  Per-role cycle counters
// (tools/microbench/bench_tc.cu) showed a single loader warp per lane quarter to be the critical
// path (load wait + split + tcgen05.st + A^T E = 1670 cycles per chunk); two halve it.
//
// Who issues the MMAs (template flag SEP):
//   SEP = false  warp 12 (a drain warp): chunk c+2 reuses the accumulator buffer of chunk c, so it is
//                issued where that warp stands once chunk c is drained.  But tcgen05.mma issue blocks on
//                the tensor pipe's short queue (12 MMAs = ~790 cycles, the math time itself), so the
//                pipe idles while warp 12 drains and polls: ~1580 cycles per chunk (measured).
//   SEP = true   a sixth warpgroup (warp 20 issues, 21-23 idle) owns the MMAs, so chunk c+1 runs while
//                chunk c is drained.  No A^T E side product in this variant (its loaders give 8
//                registers per thread to the new warpgroup): the fused DGM path forms grad[U | b]
//                elsewhere (input_map_adj) and passes PE = nullptr.  Warp 21 of that warpgroup feeds the
//                stagers: cp.async.bulk (TMA engine) of the next raw [32 x 128] S tiles into a 4-deep
//                shared-memory ring.  With register prefetch the stagers could keep only one chunk
//                (16 KB per SM) in flight and their period equalled the loaded HBM latency (~1350
//                cycles per chunk, measured) -- the critical path once the issuer was decoupled.
// grid = (Kd/128, N/128, splits).
#pragma once
#include "dgmk_gemm_tc_tn.cuh"

namespace dgmk {
namespace wg {

using tctn::KC; using tctn::BM; using tctn::BN; using tctn::HALF; using tctn::TN_OPER_BYTES; using tctn::TN_SBO;

#ifdef DGMK_WG_DEBUG
__device__ long long g_wg_prof[32];
#define WG_T(var) long long var = clock64()
#define WG_ADD(slot, t0) wg_prof[slot] += clock64() - (t0)
#define WG_DECL long long wg_prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define WG_OUT(base) if (blockIdx.x + blockIdx.y + blockIdx.z == 0 && lane == 0) { for (int q_ = 0; q_ < 8; ++q_) g_wg_prof[(base) + q_] = wg_prof[q_]; }
#else
#define WG_T(var)
#define WG_ADD(slot, t0)
#define WG_DECL
#define WG_OUT(base)
#endif
constexpr int NT = 20 * 32;                        // SEP = false
constexpr int NT_SEP = 24 * 32;                    // SEP = true
constexpr int W_STAGE = 8, W_DRAIN = 12, W_ISSUE = 20;
constexpr int HR = KC / 2;                         // rows of a chunk per loader warp
constexpr int REGS_LOAD = 72, REGS_STAGE = 96, REGS_DRAIN = 120;   // 256*72 + 128*96 + 256*120 = 640*96
// SEP: launch allocation 768 x 80 = 61440 = 256*64 + 128*96 + 256*112 + 128*32
constexpr int REGS_LOAD_SEP = 64, REGS_DRAIN_SEP = 112, REGS_ISSUE_SEP = 32;
#ifndef DGMK_WG_ISSUERS
#define DGMK_WG_ISSUERS 1   // 2: a second issuer warp on alternate chunks -- measured in the step: 18.7 -> 18.3-18.7 ms, within the noise
#endif
constexpr int NB = 3;                              // S ring stages
constexpr int STAGE_BYTES = 2 * TN_OPER_BYTES;     // hi | lo
constexpr int BAR_OFF = NB * STAGE_BYTES;
constexpr int E_OFF = BAR_OFF + 256;               // per loader warp: 32 rows of E (float4)
constexpr int SMEM_BYTES = E_OFF + 8 * HR * 16 + 1024;   // + alignment slack
constexpr int RAW_STAGES = 4;                      // SEP: raw S tiles (bulk copies), [32 rows][128] FP32
constexpr int RAW_BYTES = KC * BN * 4;
constexpr int RAW_OFF = E_OFF + 8 * HR * 16;
constexpr int SMEM_BYTES_SEP = RAW_OFF + RAW_STAGES * RAW_BYTES + 1024;
constexpr int TM_A = 0;                            // 2 buffers x (32 hi | 32 lo)
constexpr int TM_ACC = 128;                        // 2 buffers x 128
constexpr int TMEM_COLS = 512;
// kind::tf32, D = F32, M = 128, N = 128, A from TMEM (K-major), B MN-major
constexpr uint32_t IDESC_TS = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(IDESC_TS), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* src) {
  asm volatile("prefetch.global.L2 [%0];\n" ::"l"(src) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, float v0, float v1, float v2, float v3, float v4, float v5, float v6,
                                         float v7) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "f"(v0),
               "f"(v1), "f"(v2), "f"(v3), "f"(v4), "f"(v5), "f"(v6), "f"(v7)
               : "memory");
}
__device__ __forceinline__ float ldg_pinned(const float* p) {   // issued where written (see dgmk_gemm_tc.cuh)
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];\n" : "=f"(v) : "l"(p));
  return v;
}

// LDA / LDS: leading dimensions of A and S, compile-time so that the 32 row loads of a chunk are one
// base register plus immediates (with run-time strides the address arithmetic alone -- ~6
// instructions per load in a warp that has a scheduler almost to itself -- cost 1700 cycles a chunk)
template <int LDA, int LDS, bool SEP>
__global__ void __launch_bounds__(SEP ? NT_SEP : NT, 1) wgrad_ws_kernel(const float* __restrict__ A,
                                                         const float* __restrict__ S,
                                                         const float* __restrict__ E, float* __restrict__ P,
                                                         float* __restrict__ PE, int N, int Kd, int64_t M,
                                                         int64_t seg_rows, int nseg) {
  extern __shared__ char smem_raw[];
  char* smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t bar0 = tctn::smem_u32(smem + BAR_OFF);
  const uint32_t A_FULL = bar0, A_EMPTY = bar0 + 16, B_FULL = bar0 + 32, B_EMPTY = bar0 + 56, T_FULL = bar0 + 80,
                 T_EMPTY = bar0 + 96, RAW_FULL = bar0 + 128, RAW_EMPTY = bar0 + 160;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 112);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int i0 = blockIdx.y * BM;   // output rows  = columns of A
  const int j0 = blockIdx.x * BN;   // output cols  = columns of S
  const int64_t rows_per_split = seg_rows * nseg;    // seg_rows is a multiple of KC
  const int64_t mb = (int64_t)blockIdx.z * rows_per_split;
  const int64_t me = (mb + rows_per_split < M) ? mb + rows_per_split : M;
  const int64_t nchunks = (me > mb) ? (me - mb + KC - 1) / KC : 0;
  const int64_t cps = seg_rows / KC;                 // chunks per segment
  const bool do_e = !SEP && (PE != nullptr) && (blockIdx.x == 0);

  if (warp == W_DRAIN) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tctn::smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      tctn::mbar_init(A_FULL + 8 * s, 256);   // every loader thread (its tcgen05.st has completed)
      tctn::mbar_init(A_EMPTY + 8 * s, 1);    // tcgen05.commit
      tctn::mbar_init(T_FULL + 8 * s, 1);     // tcgen05.commit
      tctn::mbar_init(T_EMPTY + 8 * s, 256);  // every drain thread
    }
    for (int s = 0; s < NB; ++s) {
      tctn::mbar_init(B_FULL + 8 * s, 4);     // one arrive per stager warp
      tctn::mbar_init(B_EMPTY + 8 * s, 1);    // tcgen05.commit
    }
    if (SEP) {
      for (int s = 0; s < RAW_STAGES; ++s) {
        tctn::mbar_init(RAW_FULL + 8 * s, 1);    // expect_tx arrive + bytes
        tctn::mbar_init(RAW_EMPTY + 8 * s, 4);   // one arrive per stager warp
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp < W_STAGE) {
    if constexpr (SEP) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(REGS_LOAD_SEP));
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(REGS_LOAD));
    // ================================ A loaders: global -> hi / lo -> TMEM ===================
    // Two register sets (even / odd chunks) keep one chunk of loads in flight while the previous
    // one is split and stored; they must stay in registers (a spilled prefetch register turns
    // every load into a full-latency round trip), hence the 8-column tcgen05.st pieces.
    const int quarter = warp & 3, part = warp >> 2;
    const float* acol = A + i0 + quarter * 32 + lane;
    const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16) + TM_A + part * HR;
    double eacc[4] = {0.0, 0.0, 0.0, 0.0};   // FP32 runs of 8 chunks (128 FMAs) are added in FP64
    float ec[4] = {0.f, 0.f, 0.f, 0.f};
    float va[HR], vb[HR];
    float4 ea = make_float4(0.f, 0.f, 0.f, 0.f), eb = ea;   // lane q < 16 holds E[m0 + part*16 + q]
    auto load = [&](float (&v)[HR], float4& e, int64_t m0) {
      m0 += part * HR;
      const float* p = acol + m0 * LDA;
      if (m0 + HR <= me) {
#pragma unroll
        for (int q = 0; q < HR; ++q) v[q] = ldg_pinned(p + q * LDA);
      } else if (m0 < me) {   // last chunk of the split: clamp (no branch around a pinned load); extra rows are zeroed later
        const int last = (int)(me - 1 - m0);
#pragma unroll
        for (int q = 0; q < HR; ++q) v[q] = ldg_pinned(p + (q < last ? q : last) * LDA);
      } else {
#pragma unroll
        for (int q = 0; q < HR; ++q) v[q] = 0.f;
      }
      if (do_e) { int64_t m = m0 + (lane & (HR - 1)); m = m < me ? m : me - 1; e = __ldg(reinterpret_cast<const float4*>(E) + m); }
    };
    WG_DECL;
    auto process = [&](float (&v)[HR], const float4& e, int64_t c) {
      const int buf = (int)(c & 1);
      const int64_t m0 = mb + c * KC + part * HR;
      const int nvalid = (int)((me - m0 < HR) ? me - m0 : HR);   // may be <= 0
#pragma unroll
      for (int q = 0; q < HR; ++q) v[q] = (q < nvalid) ? v[q] : 0.f;   // rows beyond the split contribute nothing
      WG_T(t1);
      tctn::mbar_wait(A_EMPTY + 8 * buf, (uint32_t)((c >> 1) & 1) ^ 1);   // MMAs of chunk c-2 have read this buffer
      WG_ADD(1, t1);
      WG_T(t2);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
      for (int g = 0; g < HR / 8; ++g) {
        float h[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) h[q] = tctn::tf32_hi(v[g * 8 + q]);
        tmem_st8(tlane + buf * 64 + g * 8, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
        tmem_st8(tlane + buf * 64 + 32 + g * 8, v[g * 8] - h[0], v[g * 8 + 1] - h[1], v[g * 8 + 2] - h[2], v[g * 8 + 3] - h[3],
                 v[g * 8 + 4] - h[4], v[g * 8 + 5] - h[5], v[g * 8 + 6] - h[6], v[g * 8 + 7] - h[7]);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      mbar_arrive(A_FULL + 8 * buf);
      WG_ADD(2, t2);
      WG_T(t3);
      if (do_e) {   // grad[U | b] = A^T E for this thread's unit; E rows go through a warp-private patch
        float4* se = reinterpret_cast<float4*>(smem + E_OFF) + warp * HR;
        __syncwarp();
        if (lane < HR) se[lane] = e;
        __syncwarp();
#pragma unroll
        for (int q = 0; q < HR; ++q) {
          const float4 ev = se[q];   // broadcast
          ec[0] = fmaf(v[q], ev.x, ec[0]);
          ec[1] = fmaf(v[q], ev.y, ec[1]);
          ec[2] = fmaf(v[q], ev.z, ec[2]);
          ec[3] = fmaf(v[q], ev.w, ec[3]);
        }
      }
      const bool seg_end = (c + 1) % cps == 0 || c + 1 == nchunks;
      if (do_e && ((c & 7) == 7 || seg_end)) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { eacc[e] += (double)ec[e]; ec[e] = 0.f; }
      }
      if (do_e && seg_end) {   // segment complete
        float* pe = PE + (((int64_t)blockIdx.z * nseg + c / cps) * 2 + part) * 4 * N + i0 + quarter * 32 + lane;
#pragma unroll
        for (int e = 0; e < 4; ++e) { pe[(int64_t)e * N] = (float)eacc[e]; eacc[e] = 0.0; }
      }
      WG_ADD(3, t3);
    };
    if (nchunks > 0) load(va, ea, mb);
    for (int64_t c = 0; c < nchunks; c += 2) {
      WG_T(t0);
      if (c + 1 < nchunks) load(vb, eb, mb + (c + 1) * KC);
      WG_ADD(0, t0);
      process(va, ea, c);
      if (c + 1 < nchunks) {
        WG_T(t0b);
        if (c + 2 < nchunks) load(va, ea, mb + (c + 2) * KC);
        WG_ADD(0, t0b);
        process(vb, eb, c + 1);
      }
    }
    if (warp == 0) { WG_OUT(0); }
  } else if (warp < W_DRAIN) {
    if constexpr (SEP) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(REGS_STAGE));   // launch allocation: 80
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(REGS_STAGE));
    // ================================ S stagers: global -> hi / lo -> shared ring ==============
    const int sw = warp - W_STAGE;                       // rows sw + 4 q of the chunk, columns lane*4 ..
    if constexpr (SEP) {
      // raw tiles arrive by bulk copy (warp 21): shared -> registers -> hi / lo operand tile
      int stage = 0, rs = 0; uint32_t use = 0, ruse = 0;
      WG_DECL;
      for (int64_t c = 0; c < nchunks; ++c) {
        const int64_t m0 = mb + c * KC;
        WG_T(t0);
        tctn::mbar_wait(RAW_FULL + 8 * rs, ruse & 1);
        WG_ADD(0, t0);
        const char* raw = smem + RAW_OFF + rs * RAW_BYTES + sw * (BN * 4) + lane * 16;
        float4 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = *reinterpret_cast<const float4*>(raw + q * 4 * (BN * 4));
        char* sh = smem + stage * STAGE_BYTES;
        WG_T(t1);
        tctn::mbar_wait(B_EMPTY + 8 * stage, (use & 1) ^ 1);
        WG_ADD(1, t1);
        WG_T(t2);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 x = (m0 + q * 4 + sw < me) ? v[q] : make_float4(0.f, 0.f, 0.f, 0.f);
          tctn::split_store(sh, sh + TN_OPER_BYTES, tctn::tn_off(q * 4 + sw, lane * 4), x);
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy stores -> UMMA
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(B_FULL + 8 * stage);
          mbar_arrive(RAW_EMPTY + 8 * rs);     // every lane has its raw values in registers
        }
        WG_ADD(2, t2);
        if (++stage == NB) { stage = 0; ++use; }
        if (++rs == RAW_STAGES) { rs = 0; ++ruse; }
      }
      if (warp == W_STAGE) { WG_OUT(8); }
    } else {
    const float* sbase = S + j0 + lane * 4;
    float4 va[8], vb[8];
    auto load = [&](float4 (&v)[8], int64_t m0) {
      const float* p = sbase + (m0 + sw) * LDS;
      if (m0 + KC <= me) {
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = tctn::ldg_f4_pinned(p + q * 4 * LDS);
      } else {
        const int last = (int)(me - 1 - m0 - sw);   // may be negative: then row me-1 itself
#pragma unroll
        for (int q = 0; q < 8; ++q) { const int r = q * 4 < last ? q * 4 : last; v[q] = tctn::ldg_f4_pinned(p + r * LDS); }
      }
    };
    int stage = 0; uint32_t use = 0;
    WG_DECL;
    auto process = [&](float4 (&v)[8], int64_t c) {
      const int64_t m0 = mb + c * KC;
      char* sh = smem + stage * STAGE_BYTES;
      WG_T(t1);
      tctn::mbar_wait(B_EMPTY + 8 * stage, (use & 1) ^ 1);
      WG_ADD(1, t1);
      WG_T(t2);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 x = (m0 + q * 4 + sw < me) ? v[q] : make_float4(0.f, 0.f, 0.f, 0.f);
        tctn::split_store(sh, sh + TN_OPER_BYTES, tctn::tn_off(q * 4 + sw, lane * 4), x);
      }
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy stores -> UMMA
      __syncwarp();
      if (lane == 0) mbar_arrive(B_FULL + 8 * stage);
      WG_ADD(2, t2);
      if (++stage == NB) { stage = 0; ++use; }
    };
    if (nchunks > 0) load(va, mb);
    for (int64_t c = 0; c < nchunks; c += 2) {
      WG_T(t0);
      if (c + 1 < nchunks) load(vb, mb + (c + 1) * KC);
      WG_ADD(0, t0);
      process(va, c);
      if (c + 1 < nchunks) {
        if (c + 2 < nchunks) load(va, mb + (c + 2) * KC);
        process(vb, c + 1);
      }
    }
    if (warp == W_STAGE) { WG_OUT(8); }
    }
  } else {
    WG_DECL;
    // warp-uniform; one elected lane issues the 12 MMAs of chunk c and the three commits
    auto issue = [&](int64_t c) {
      const int buf = (int)(c & 1);
      const int stage = (int)(c % NB);
      const uint32_t ph = (uint32_t)((c >> 1) & 1);
      WG_T(t0);
      tctn::mbar_wait(T_EMPTY + 8 * buf, ph ^ 1);    // accumulator buffer drained by all 8 warps
      WG_ADD(2, t0);
      WG_T(t1);
      tctn::mbar_wait(A_FULL + 8 * buf, ph);
      WG_ADD(3, t1);
      WG_T(t2);
      tctn::mbar_wait(B_FULL + 8 * stage, (uint32_t)((c / NB) & 1));
      WG_ADD(4, t2);
      WG_T(t3);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint32_t sb = tctn::smem_u32(smem + stage * STAGE_BYTES);
      const uint64_t dBh = tctn::make_desc_mn(sb), dBl = tctn::make_desc_mn(sb + TN_OPER_BYTES);
      const uint32_t ah = tmem + TM_A + buf * 64, al = ah + 32;
      const uint32_t d = tmem + TM_ACC + buf * BN;
      if (tctn::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < KC / 8; ++ks) {   // small terms first; a k-step of 8 rows = two 4-row groups
          const uint64_t adv = (uint64_t)((ks * 2 * TN_SBO) >> 4);
          mma_ts(d, al + ks * 8, dBh + adv, ks > 0 ? 1u : 0u);
          mma_ts(d, ah + ks * 8, dBl + adv, 1u);
        }
#pragma unroll
        for (int ks = 0; ks < KC / 8; ++ks) {
          const uint64_t adv = (uint64_t)((ks * 2 * TN_SBO) >> 4);
          mma_ts(d, ah + ks * 8, dBh + adv, 1u);
        }
        tctn::mma_commit(A_EMPTY + 8 * buf);
        tctn::mma_commit(B_EMPTY + 8 * stage);
        tctn::mma_commit(T_FULL + 8 * buf);
      }
      __syncwarp();
      WG_ADD(5, t3);
    };
    if (warp < W_ISSUE) {
      if constexpr (SEP) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(REGS_DRAIN_SEP));
      else asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(REGS_DRAIN));
      // ============================== drain: chunk results -> RN registers (!SEP: warp 12 also issues) ===
      const int quarter = warp & 3, half = (warp - W_DRAIN) >> 2;
      float acc[HALF];
#pragma unroll
      for (int j = 0; j < HALF; ++j) acc[j] = 0.f;
      if (!SEP && warp == W_DRAIN) {
        if (nchunks > 0) issue(0);
        if (nchunks > 1) issue(1);
      }
      for (int64_t c = 0; c < nchunks; ++c) {
        const int buf = (int)(c & 1);
        WG_T(t0);
        tctn::mbar_wait(T_FULL + 8 * buf, (uint32_t)((c >> 1) & 1));
        WG_ADD(0, t0);
        WG_T(t1);
        tctn::drain_half(tmem + TM_ACC, buf, quarter, half, acc);   // fences inside
        mbar_arrive(T_EMPTY + 8 * buf);
        WG_ADD(1, t1);
        if (!SEP && warp == W_DRAIN && c + 2 < nchunks) issue(c + 2);
        if ((c + 1) % cps == 0 || c + 1 == nchunks) {   // segment complete: this thread owns output row
          // i0 + quarter*32 + lane, columns j0 + half*64 .. of the segment's partial
          float* prow = P + ((int64_t)blockIdx.z * nseg + c / cps) * N * Kd + (int64_t)(i0 + quarter * 32 + lane) * Kd + j0 + half * HALF;
#pragma unroll
          for (int q = 0; q < HALF / 4; ++q) {
            *reinterpret_cast<float4*>(prow + q * 4) = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
            acc[4 * q] = 0.f; acc[4 * q + 1] = 0.f; acc[4 * q + 2] = 0.f; acc[4 * q + 3] = 0.f;
          }
        }
      }
#ifdef DGMK_WG_DEBUG
      wg_prof[6] = nchunks;
#endif
      if (warp == W_DRAIN) { WG_OUT(24); }
    } else {
      // ============================== SEP: MMA issuer (warp 20; 21-23 only hand their registers over) ===
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(REGS_ISSUE_SEP));
      if (warp == W_ISSUE || (DGMK_WG_ISSUERS == 2 && warp == W_ISSUE + 2)) {
        // (DGMK_WG_ISSUERS == 2: issuers on alternate chunks, issuer i owns A buffer i and accumulator buffer i.  It pays
        // in the data-gradient and lane kernels, not here: loaders and stagers, not the issuer, set this kernel's pace)
#pragma unroll 1
        for (int64_t c = (warp - W_ISSUE) >> 1; c < nchunks; c += DGMK_WG_ISSUERS) issue(c);
        if (warp == W_ISSUE) { WG_OUT(16); }
      } else if (warp == W_ISSUE + 1) {
        // bulk-copy producer: raw S tile of chunk c -> ring stage (rows of S are contiguous when LDS == 128)
        int rs = 0; uint32_t ruse = 0;
#pragma unroll 1
        for (int64_t c = 0; c < nchunks; ++c) {
          const int64_t m0 = mb + c * KC;
          const int nrows = (int)((me - m0 < KC) ? me - m0 : KC);
          tctn::mbar_wait(RAW_EMPTY + 8 * rs, (ruse & 1) ^ 1);
          const uint32_t dst = tctn::smem_u32(smem + RAW_OFF + rs * RAW_BYTES);
          if (lane == 0) mbar_expect_tx(RAW_FULL + 8 * rs, (uint32_t)nrows * BN * 4);
          __syncwarp();
          if (LDS == BN) {
            if (lane == 0) bulk_g2s(dst, S + m0 * LDS + j0, (uint32_t)nrows * BN * 4, RAW_FULL + 8 * rs);
          } else if (lane < nrows) {
            bulk_g2s(dst + lane * (BN * 4), S + (m0 + lane) * LDS + j0, BN * 4, RAW_FULL + 8 * rs);
          }
          // the same chunk of A into L2: this warp runs several chunks ahead of the loaders (ring depth),
          // whose register prefetch covers only one chunk -- their loads then see L2, not HBM, latency
          // (prefetch.global.L2 through the LSU, one 128-byte line per lane: cp.async.bulk.prefetch per row cost the
          // TMA engine more than the copies themselves -- the stagers then waited 1200 cycles per chunk)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int r = q * 8 + (lane >> 2);
            if (r < nrows) prefetch_l2(A + (m0 + r) * LDA + i0 + (lane & 3) * 32);
          }
          if (++rs == RAW_STAGES) { rs = 0; ++ruse; }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == W_DRAIN) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

}  // namespace wg
}  // namespace dgmk
