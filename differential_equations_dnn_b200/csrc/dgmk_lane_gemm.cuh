// Warp-specialised persistent tcgen05 GEMM with the hidden UNITS on the TMEM lanes and the
// collocation ROWS on the TMEM columns (sm_100a), 3xTF32 split, K = 128:
//
//   D[j, r] = sum_k W[g*128 + j, k] * X[r, k]        j = 0..127 (unit), g = gate of this CTA
//
// i.e. the transpose of what dgmk_gemm_tc.cuh computes.  Two reasons for this orientation:
//
// 1. One epilogue thread owns ONE hidden unit j and sees every channel of a collocation point
//    (consecutive rows = consecutive accumulator columns) in its own registers, which is exactly
//    the shape of the element-wise jet stages (dgmk_ops.h: one (point, unit) pair, all channels).
//    Gate activations, s*R, the state update and their adjoints therefore run as the EPILOGUE of
//    the GEMM that feeds them, straight out of TMEM, instead of as separate kernels that re-read
//    the GEMM output from HBM; the epilogue's global accesses are 128 contiguous bytes per warp
//    and row.
// 2. The weights become the M-side operand, which tcgen05.mma can read from TENSOR MEMORY.
//    Measured (tools/microbench/bench_mma_rate2.cu, bench_lane.cu): a kind::tf32 MMA with both
//    operands in shared memory re-reads the 4 KB M-side slab for every K=8 step and is bound by
//    the 128 B/clk shared-memory pipe (N=64: 48 cycles, N=128: 64, against a math floor of 32 /
//    64), and in a real kernel that pipe is shared with the operand staging and the epilogue's
//    global traffic -- the streaming tile and a first smem-weights version of this kernel both
//    stalled at ~105 TFLOP/s with the MMA warp back-pressured at ~90 cycles per N=64 MMA.  With
//    the CTA's 128 x 128 weight block (tf32 hi and lo) resident in 256 TMEM columns, an MMA reads
//    only the 2 KB row slab from shared memory.
//
// Arithmetic is the one of dgmk_gemm_tc.cuh (read that header first): per K chunk of 32,
// lo*hi + hi*lo then hi*hi from zero in TMEM, the four chunk results summed in round-to-nearest
// registers.  Each chunk has its own 64 accumulator columns: TMEM = 256 (weights) + 4 x 64.
//
// One CTA per SM, 24 warps (6 warpgroups), decoupled by mbarriers:
//   warp 0       bulk-copy producer: cp.async.bulk (TMA engine; rows are contiguous, no tensor map
//                needed) of the next [64 x 128] FP32 row tile into a 3-deep raw ring
//   warps 4-7    transform: raw tile -> tf32 hi / lo -> UMMA canonical K-major operand tile (2-deep)
//   warps 1-2    MMA issuers on alternate K chunks (round 2: the plain GEMM 0.37 -> 0.28 ms per 524 288 rows x 384
//                units; the fused stages are epilogue- / HBM-bound and gain little): per K chunk 12 tcgen05.mma
//                (A = weights in TMEM, B = rows in smem, M=128, N=64, K=8) -> tcgen05.commit per chunk; each frees
//                the operand tile after its last chunk
//   warps 8-23   epilogue (warp w: TMEM lanes 32*(w%4).., 16 of the tile's 64 rows): tcgen05.ld each
//                chunk result as soon as it is complete (handing its columns straight back to the
//                MMA warp), add, then run the fused stage (EPI); setmaxnreg moves the producers'
//                spare registers to these four warpgroups
// CTA b works on gate b % ngates for its whole life (weight-stationary); the CTAs of the different
// gates walk the same row tiles at the same time and share them in L2.
#pragma once
#include "dgmk_gemm_tc.cuh"

namespace dgmk {
namespace lg {

#ifdef DGMK_LG_DEBUG   // microbenchmark-only switches: bit 0 = skip the MMAs, bit 1 = skip the transform stores
__device__ int g_lg_dbg = 0;
__device__ int g_lg_prof_cta = 0;
__device__ long long g_lg_prof[32];
#define LG_DBG(bit) (g_lg_dbg & (bit))
#define LG_T(var) long long var = clock64()
#define LG_ADD(slot, t0) lg_prof[slot] += clock64() - (t0)
#define LG_PROF_DECL long long lg_prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define LG_PROF_OUT(base) if (blockIdx.x == g_lg_prof_cta && lane == 0) { for (int q_ = 0; q_ < 8; ++q_) g_lg_prof[(base) + q_] = lg_prof[q_]; }
#else
#define LG_DBG(bit) 0
#define LG_T(var)
#define LG_ADD(slot, t0)
#define LG_PROF_DECL
#define LG_PROF_OUT(base)
#endif

constexpr int NU = 128;    // units per CTA (UMMA M, TMEM lanes)
constexpr int NR = 64;     // rows per tile (UMMA N, TMEM columns per chunk)
constexpr int KC = 32;     // K per chunk
constexpr int NCH = 4;     // K = 128
constexpr int KTOT = KC * NCH;
constexpr int X_LBO = 1024 + 16;   // row operand: padded so the transform stores are conflict-free
constexpr int SBO = 128;
constexpr int X_OPER = 8 * X_LBO;  // [64 rows x 32 k]
constexpr int RAW_STAGES = 3;
constexpr int RAW_BYTES = NR * KTOT * 4;
constexpr int X_STAGES = 2;                          // whole tiles
constexpr int X_STAGE_BYTES = NCH * 2 * X_OPER;      // per chunk: hi | lo
constexpr int RAW_OFF = 0;
constexpr int X_OFF = RAW_OFF + RAW_STAGES * RAW_BYTES;
constexpr int BAR_OFF = X_OFF + X_STAGES * X_STAGE_BYTES;
constexpr int SMEM_BYTES = BAR_OFF + 192;
constexpr int NTW = 4;                // transform warps
// warp roles, aligned to warpgroups so that registers can be moved between them (setmaxnreg):
//   WG0: warp 0 copy, warp 1 MMA, warps 2-3 idle | WG1: warps 4-7 transform | WG2-3: warps 8-15 epilogue
constexpr int W_TRANSFORM = 4, W_EPI = 8;
constexpr int EPI_WARPS = 16;                        // 4 per scheduler: the fused stages are latency-bound
constexpr int EPI_ROWS = NR / (EPI_WARPS / 4);       // accumulator columns (= rows) per epilogue thread
constexpr int EPI_GROUPS = EPI_ROWS / 8;
constexpr int NT = (W_EPI + EPI_WARPS) * 32;
constexpr int REGS_PRODUCER = 32, REGS_EPI = 104;    // must fit the launch allocation (768 x 80 = 61440): 256 x 32 + 512 x 104 = 61440
constexpr int NCONS = EPI_WARPS * 32;
#ifndef DGMK_LG_ISSUERS
#define DGMK_LG_ISSUERS 2
#endif
constexpr int NISS = DGMK_LG_ISSUERS;   // MMA issuer warps (warps 1 .. NISS), chunk c belongs to issuer c % NISS
constexpr int TMEM_COLS = 512;
constexpr int TM_WHI = 0, TM_WLO = KTOT, TM_ACC = 2 * KTOT;   // TMEM column map

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// kind::tf32, D = F32, A/B = TF32 both K-major, M = 128, N = 64
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NR >> 3) << 17) | ((uint32_t)(NU >> 4) << 24);

// A (weights) from tensor memory, B (rows) from shared memory
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(IDESC), "r"(accumulate)
      : "memory");
}
// one lane of a converged warp (the tcgen05 issue idiom: the surrounding loop stays warp-uniform,
// so descriptors and addresses live in uniform registers instead of being broadcast per MMA)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
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
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float4 (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "f"(v[0].x), "f"(v[0].y), "f"(v[0].z), "f"(v[0].w), "f"(v[1].x), "f"(v[1].y), "f"(v[1].z), "f"(v[1].w),
      "f"(v[2].x), "f"(v[2].y), "f"(v[2].z), "f"(v[2].w), "f"(v[3].x), "f"(v[3].y), "f"(v[3].z), "f"(v[3].w),
      "f"(v[4].x), "f"(v[4].y), "f"(v[4].z), "f"(v[4].w), "f"(v[5].x), "f"(v[5].y), "f"(v[5].z), "f"(v[5].w),
      "f"(v[6].x), "f"(v[6].y), "f"(v[6].z), "f"(v[6].w), "f"(v[7].x), "f"(v[7].y), "f"(v[7].z), "f"(v[7].w)
      : "memory");
}

template <bool B> struct FullTag { static constexpr bool value = B; };

// ---- plain epilogue: C[r, gate*128 + j] (+)= D[j, r]  (interface: see dgmk_lane_epi.cuh) --------
template <bool ACCUM>
struct StoreEpi {
  float* C; int64_t ldc;
  struct Const { int col; };
  struct Tile {};
  struct State {};
  struct Pre { float old[8]; };
  __device__ __forceinline__ void finish(const State&, const Const&, int) const {}
  __device__ __forceinline__ Const init(int gate, int j) const { Const k; k.col = gate * NU + j; return k; }
  __device__ __forceinline__ void tile(Tile&, const Const&, int64_t, int64_t, int) const {}
  template <bool FULL>
  __device__ __forceinline__ void prefetch(Pre& p, const Tile&, const Const& k, int64_t row0, int, int64_t M) const {
    if (ACCUM) {
#pragma unroll
      for (int q = 0; q < 8; ++q) p.old[q] = (FULL || row0 + q < M) ? C[(row0 + q) * ldc + k.col] : 0.f;
    }
  }
  // a[q] = D[j, row0 + q], q = 0..7; rows >= M are padding
  template <bool FULL>
  __device__ __forceinline__ void apply(const Pre& p, State&, const Tile&, int, const Const& k, int64_t row0, int64_t M, const float (&a)[8]) const {
    float* c = C + row0 * ldc + k.col;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (FULL || row0 + q < M) c[q * ldc] = ACCUM ? p.old[q] + a[q] : a[q];
  }
};

// X: [M, 128] rows with leading dimension ldx; Wt: [ngates*128, 128] K-major with its tf32 hi / lo
// copies hl_stride / 2*hl_stride further.  CTAs [0, c0) work on gate 0, [c0, c0+c1) on gate 1, the
// rest on gate 2 (ngates == 3) -- the gates need not get the same number of CTAs: the R gate of a
// DGM layer also forms s*R and costs ~15 % more per tile (measured), so it gets more of them and
// all three sweep the rows at the same pace (which also keeps the shared row tiles in L2).
template <class EPI>
__global__ void __launch_bounds__(NT, 1) lane_gemm_kernel(const float* __restrict__ X, int64_t ldx,
                                                          const float* __restrict__ Wt, int64_t ldw, int64_t hl_stride,
                                                          int64_t M, int ngates, int c0, int c1, const EPI epi) {
  extern __shared__ __align__(128) char smem[];
  const uint32_t bar0 = tc::smem_u32(smem + BAR_OFF);
  const uint32_t RAW_FULL = bar0, RAW_EMPTY = bar0 + 24, OP_FULL = bar0 + 48, OP_EMPTY = bar0 + 64, TC_FULL = bar0 + 80,
                 TC_EMPTY = bar0 + 112;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 144);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;
  const int gate = (ngates == 1 || b < c0) ? 0 : (b < c0 + c1 ? 1 : 2);
  const int64_t grp = (gate == 0) ? b : (gate == 1 ? b - c0 : b - c0 - c1);
  const int64_t ngrp = (ngates == 1) ? gridDim.x : (gate == 0 ? c0 : (gate == 1 ? c1 : (int)gridDim.x - c0 - c1));
  const int64_t ntiles = (M + NR - 1) / NR;

  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tc::smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < RAW_STAGES; ++s) {
      tc::mbar_init(RAW_FULL + 8 * s, 1);      // expect_tx arrive + bytes
      tc::mbar_init(RAW_EMPTY + 8 * s, NTW);   // one arrive per transform warp
    }
    for (int s = 0; s < X_STAGES; ++s) {
      tc::mbar_init(OP_FULL + 8 * s, NTW);
      tc::mbar_init(OP_EMPTY + 8 * s, NISS);   // tcgen05.commit of every issuer
    }
    for (int c = 0; c < NCH; ++c) {
      tc::mbar_init(TC_FULL + 8 * c, 1);       // tcgen05.commit
      tc::mbar_init(TC_EMPTY + 8 * c, NCONS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp >= W_EPI && warp < W_EPI + 8) {
    // resident weights of this gate -> TMEM: lane = unit, column = k (32-bit cells); the first four
    // epilogue warps write the tf32-hi copy, the other four the lo copy
    const int quarter = warp & 3, half = (warp - W_EPI) >> 2;
    const float* src = Wt + (1 + half) * hl_stride + (int64_t)(gate * NU + quarter * 32 + lane) * ldw;
    const uint32_t tw = tmem + ((uint32_t)(quarter * 32) << 16) + (half ? TM_WLO : TM_WHI);
#pragma unroll 1
    for (int kb = 0; kb < KTOT / 32; ++kb) {
      float4 v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = __ldg(reinterpret_cast<const float4*>(src + kb * 32) + q);
      tmem_st32(tw + kb * 32, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

  // The epilogue warps hold the prefetch ring and the accumulators: setmaxnreg (one per warpgroup,
  // at the top of its role branch so that ptxas allocates each branch against its own limit) gives
  // them the registers the copy / MMA / transform warps do not need.
  {
   if (warp < W_TRANSFORM) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(REGS_PRODUCER));
    if (warp == 0) {
      // ================================ bulk-copy producer ================================
      uint32_t i = 0, rs = 0, use = 0;
      LG_PROF_DECL;
      LG_T(tstart);
      for (int64_t t = grp; t < ntiles; t += ngrp, ++i) {
        LG_T(t0);
        tc::mbar_wait(RAW_EMPTY + 8 * rs, (use & 1) ^ 1);
        LG_ADD(0, t0);
        const int64_t row0 = t * NR;
        const int nrows = (int)((M - row0 < NR) ? M - row0 : NR);
        const uint32_t dst = tc::smem_u32(smem + RAW_OFF + rs * RAW_BYTES);
        if (lane == 0) mbar_expect_tx(RAW_FULL + 8 * rs, (uint32_t)nrows * KTOT * 4);
        __syncwarp();
        if (ldx == KTOT) {
          if (lane == 0) bulk_g2s(dst, X + row0 * ldx, (uint32_t)nrows * KTOT * 4, RAW_FULL + 8 * rs);
        } else {
          for (int r = lane; r < nrows; r += 32) bulk_g2s(dst + r * KTOT * 4, X + (row0 + r) * ldx, KTOT * 4, RAW_FULL + 8 * rs);
        }
        if (++rs == RAW_STAGES) { rs = 0; ++use; }
      }
      LG_ADD(1, tstart);
#ifdef DGMK_LG_DEBUG
      lg_prof[2] = i;
#endif
      LG_PROF_OUT(0);
    } else if (warp <= NISS) {
      // ================================ MMA issuer(s) ======================================
      // the whole warp runs the loop (uniform control flow); one elected lane issues
      uint32_t i = 0;
      LG_PROF_DECL;
      for (int64_t t = grp; t < ntiles; t += ngrp, ++i) {
        const uint32_t ts = i & 1;
        LG_T(t1);
        tc::mbar_wait(OP_FULL + 8 * ts, (i >> 1) & 1);
        LG_ADD(1, t1);
        const uint32_t xb = tc::smem_u32(smem + X_OFF + ts * X_STAGE_BYTES);
#pragma unroll 1   // this warp lives on 40 registers (setmaxnreg): one chunk's descriptors at a time
        for (int c = warp - 1; c < NCH; c += NISS) {
          LG_T(t0);
          tc::mbar_wait(TC_EMPTY + 8 * c, (i & 1) ^ 1);   // epilogue has read the previous tile's chunk c
          LG_ADD(0, t0);
          LG_T(t2);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const uint64_t dXh = make_desc(xb + c * 2 * X_OPER, X_LBO), dXl = make_desc(xb + c * 2 * X_OPER + X_OPER, X_LBO);
          const uint32_t d = tmem + TM_ACC + c * NR;
          const uint32_t wh = tmem + TM_WHI + c * KC, wl = tmem + TM_WLO + c * KC;
          if (elect_one()) {
            if (!LG_DBG(1)) {
#pragma unroll
              for (int ks = 0; ks < KC / 8; ++ks) {   // small terms first
                const uint64_t ax = (uint64_t)((ks * 2 * X_LBO) >> 4);
                mma_ts(d, wh + ks * 8, dXl + ax, ks > 0 ? 1u : 0u);
                mma_ts(d, wl + ks * 8, dXh + ax, 1u);
              }
#pragma unroll
              for (int ks = 0; ks < KC / 8; ++ks) {
                const uint64_t ax = (uint64_t)((ks * 2 * X_LBO) >> 4);
                mma_ts(d, wh + ks * 8, dXh + ax, 1u);
              }
            }
            tc::mma_commit(TC_FULL + 8 * c);                        // chunk result complete
            if (c + NISS >= NCH) tc::mma_commit(OP_EMPTY + 8 * ts);   // operand tile reusable once all MMAs (of every issuer) have read it
          }
          __syncwarp();
          LG_ADD(2, t2);
        }
      }
      if (warp == 1) { LG_PROF_OUT(8); }
    }
   } else if (warp < W_EPI) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(REGS_PRODUCER));
    {
      // ================================ transform ==========================================
      const int ttid = tid - W_TRANSFORM * 32;
      const int piece = ttid & 7, rbase = ttid >> 3;   // rows rbase + 16 q
      const int st_off = piece * X_LBO + (rbase >> 3) * SBO + (rbase & 7) * 16;   // + q * 2 * SBO
      uint32_t i = 0, rs = 0, use = 0;
      LG_PROF_DECL;
      for (int64_t t = grp; t < ntiles; t += ngrp, ++i) {
        const uint32_t ts = i & 1;
        LG_T(t0);
        tc::mbar_wait(RAW_FULL + 8 * rs, use & 1);
        LG_ADD(0, t0);
        LG_T(t1);
        tc::mbar_wait(OP_EMPTY + 8 * ts, ((i >> 1) & 1) ^ 1);
        LG_ADD(1, t1);
        LG_T(t2);
        const char* raw = smem + RAW_OFF + rs * RAW_BYTES + rbase * (KTOT * 4) + piece * 16;
        char* xt = smem + X_OFF + ts * X_STAGE_BYTES;
#pragma unroll 1   // 40 registers per thread here (setmaxnreg): four 16-byte pieces in flight
        for (int c = 0; c < NCH; ++c) {
          float4 v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) v[q] = *reinterpret_cast<const float4*>(raw + q * 16 * (KTOT * 4) + c * (KC * 4));
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (!LG_DBG(2)) tc::split_store(xt + c * 2 * X_OPER, xt + c * 2 * X_OPER + X_OPER, st_off + q * 2 * SBO, v[q]);
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy stores -> UMMA
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(OP_FULL + 8 * ts);
          mbar_arrive(RAW_EMPTY + 8 * rs);
        }
        LG_ADD(2, t2);
        if (++rs == RAW_STAGES) { rs = 0; ++use; }
      }
      if (warp == W_TRANSFORM) { LG_PROF_OUT(16); }
    }
   } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(REGS_EPI));
    {
      // ================================ epilogue ===========================================
      const int quarter = warp & 3, part = (warp - W_EPI) >> 2;
      const int j = quarter * 32 + lane;
      const uint32_t tbase = tmem + ((uint32_t)(quarter * 32) << 16) + TM_ACC + part * EPI_ROWS;
      const typename EPI::Const ek = epi.init(gate, j);
      // The global loads of a group (8 rows) are issued one group ahead of their use, across tile
      // boundaries: pre[cg] belongs to group cg of the current tile; while group 0 is computed the
      // loads of group 1 go out, while group 1 is computed those of the next tile's group 0.
      typename EPI::Tile et, et_next;
      typename EPI::Pre pre[EPI_GROUPS];
      typename EPI::State est{};   // per-thread state that lives across tiles (e.g. gradient partial sums)
      uint32_t i = 0;
      LG_PROF_DECL;
      if (grp < ntiles) {
        const int64_t r0 = grp * NR + part * EPI_ROWS;
        epi.tile(et, ek, r0, M, lane);
        epi.template prefetch<false>(pre[0], et, ek, r0, 0, M);
      }
      for (int64_t t = grp; t < ntiles; t += ngrp, ++i) {
        const int64_t row0 = t * NR + part * EPI_ROWS;
        const int64_t nrow0 = (t + ngrp) * NR + part * EPI_ROWS;
        const bool has_next = t + ngrp < ntiles;
        // the next tile's point coordinates: fetched a whole tile before their first use (a shuffle)
        if (has_next) epi.tile(et_next, ek, nrow0, M, lane);
        float a[EPI_ROWS];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          LG_T(t0);
          tc::mbar_wait(TC_FULL + 8 * c, i & 1);
          LG_ADD(0, t0);
          LG_T(t1);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          uint32_t v[EPI_GROUPS][8];
#pragma unroll
          for (int g = 0; g < EPI_GROUPS; ++g) tmem_ld8(tbase + c * NR + g * 8, v[g]);
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
          mbar_arrive(TC_EMPTY + 8 * c);   // these columns can take the next tile's chunk
#pragma unroll
          for (int g = 0; g < EPI_GROUPS; ++g)
#pragma unroll
            for (int q = 0; q < 8; ++q)
              a[g * 8 + q] = (c == 0) ? __uint_as_float(v[g][q]) : a[g * 8 + q] + __uint_as_float(v[g][q]);
          LG_ADD(1, t1);
        }
        LG_T(t2);
        // fast path: this part of the tile and of the next one lie inside [0, M): no clamps, no predicates
        auto groups = [&](auto full_tag) {
          constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
          for (int cg = 0; cg < EPI_GROUPS; ++cg) {
            if (cg + 1 < EPI_GROUPS) {
              epi.template prefetch<FULL>(pre[cg + 1], et, ek, row0 + (cg + 1) * 8, cg + 1, M);
            } else if (has_next) {
              epi.template prefetch<FULL>(pre[0], et_next, ek, nrow0, 0, M);
            }
            const float a8[8] = {a[cg * 8], a[cg * 8 + 1], a[cg * 8 + 2], a[cg * 8 + 3], a[cg * 8 + 4], a[cg * 8 + 5], a[cg * 8 + 6], a[cg * 8 + 7]};
            epi.template apply<FULL>(pre[cg], est, et, cg, ek, row0 + cg * 8, M, a8);
          }
        };
        if (row0 + EPI_ROWS <= M && (!has_next || nrow0 + EPI_ROWS <= M)) groups(FullTag<true>{});
        else groups(FullTag<false>{});
        et = et_next;
        LG_ADD(2, t2);
      }
      epi.finish(est, ek, (int)blockIdx.x * (EPI_WARPS / 4) + part);
      if (warp == W_EPI) { LG_PROF_OUT(24); }
    }
   }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

}  // namespace lg
}  // namespace dgmk
