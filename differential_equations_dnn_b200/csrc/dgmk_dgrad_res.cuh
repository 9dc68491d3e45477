// Weight-resident tcgen05 data-gradient kernel (sm_100a), 3xTF32 split, for the K = 3H contraction of a DGM layer's
// reverse pass (dgm_net.py:53-68 differentiated: s bar += [abar_Z | abar_G | abar_R] [W_z; W_g; W_r]):
//
//   C[M, 128] (+)= A[M, K] * Bt[128, K]^T        K = 32 * nch <= 384, A / C row-major FP32, Bt K-major with its tf32
//                                                hi / lo copies hl_stride / 2 * hl_stride further
//
// The round-1 streaming tile (dgmk_gemm_tc.cuh) was bound by neither HBM (0.43 of the measured peak) nor the tensor
// pipe (0.46 of the 3xTF32 ceiling): every 128-row tile re-streamed the 393 KB of split weights from L2 and staged BOTH
// operands through shared memory, whose 128 B/clk pipe the SS-form MMA also reads 8 KB per K = 8 step from
// (profiles/r01_notes.md section 12: a better pipeline alone did not help; fewer weight bytes per row is the lever).
// Here neither operand is staged per tile:
//
//   * the WEIGHTS stay in shared memory for the whole life of a persistent CTA.  hi + lo of all 128 output units do not
//     fit (393 KB), so a CTA owns HALF of the output units: [64 x K] hi | lo = 196 608 B in the UMMA canonical K-major
//     no-swizzle layout, one 8 KB operand per K chunk of 32.  CTAs 2p and 2p + 1 walk the same 128-row tiles at the
//     same time (the second read of a row tile is an L2 hit) and write the two 64-column halves of the result.
//   * the ROWS are the M-side operand and live in TENSOR MEMORY (TS form): thread r splits the 128 bytes of row r of the
//     chunk into tf32 hi / lo in registers and tcgen05.st's its own TMEM lane; the chunk reaches it through a two-stage
//     raw ring filled by TMA tensor copies, and an MMA reads only the 2 KB weight slab from shared memory.  Measured on the way (tools/microbench/bench_dgrad.cu): row-per-thread global loads touch 32 lines per
//     warp instruction and are L1-wavefront bound (1.8 ms per 2^21 rows against 1.65 for the streaming tile); coalesced
//     register loads + a warp-private transposition patch stall on the load queue (2.6 ms); 16-byte cp.async's from
//     two warps block at issue (1.7-1.9 ms).
//
// Arithmetic is the one of dgmk_gemm_tc.cuh: per K chunk lo*hi + hi*lo then hi*hi from zero in TMEM (12 MMAs, M = 128,
// N = 64, K = 8), the chunk results added in round-to-nearest FP32 registers (the tensor core adds with truncation).
//
// 20 warps, decoupled by mbarriers.  The unit of hand-over between the roles is a ROUND of two chunks (a hop through a
// barrier costs every role a few hundred cycles -- more than the work of one chunk; with one hop per chunk the bare
// pipeline, no MMAs, ran at ~780 cycles per chunk against 640 cycles of MMAs).  TMEM: 2 round slots x [2 x (32 hi | 32 lo)
// columns of A] + 2 round slots x [2 x 64 accumulator columns].
//   warp 17      copy: one TMA tensor copy (cp.async.bulk.tensor.2d, 128-byte swizzle) of chunk g into raw-ring stage
//                g % 2 (all the shared memory the weights leave: 2 x 16 KB), L2 prefetch of the chunk 8 ahead
//   warps 0-7    transformers: warpgroup w takes chunk 2 R + w of round R (thread = row = TMEM lane)
//   warps 8-15   drain: warp w reads lanes 32 (w % 4).., columns 32 ((w - 8) / 4).. of the round's two chunk results and
//                hands the columns back; the 16x256b fragment shape of tcgen05.ld puts 32 contiguous bytes of a row of C
//                into four neighbouring threads, so the read-modify-write of C after the last round runs on whole
//                sectors without a transposition (no shared memory left for one: weights + raw ring)
//   warps 16, 18 MMA issuers on alternate rounds (one elected lane each, warp-uniform control flow); warp 19 only hands
//                its registers over
#pragma once
#include <cuda.h>   // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)
#include "dgmk_gemm_tc.cuh"
#include "dgmk_lane_gemm.cuh"

namespace dgmk {
namespace dg {

#ifdef DGMK_DG_DEBUG   // microbenchmark-only switches (tools/microbench/bench_dgrad.cu)
__device__ int g_dg_dbg = 0;       // bit 0: skip the MMAs, bit 1: skip the global loads, bit 2: skip the C read-modify-write, bit 3: no L2 prefetch
#define DG_DBG(bit) (g_dg_dbg & (bit))
#else
#define DG_DBG(bit) 0
#endif
#ifdef DGMK_DG_PROF    // microbenchmark-only: per-role cycle counters of CTA 0
__device__ long long g_dg_prof[32];
#define DG_T(var) long long var = clock64()
#define DG_ADD(slot, t0) dg_prof[slot] += clock64() - (t0)
#define DG_DECL long long dg_prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define DG_OUT(base) if (blockIdx.x == 0 && lane == 0) { for (int q_ = 0; q_ < 8; ++q_) g_dg_prof[(base) + q_] = dg_prof[q_]; }
#else
#define DG_T(var)
#define DG_ADD(slot, t0)
#define DG_DECL
#define DG_OUT(base)
#endif

constexpr int BM = 128;            // rows per tile (UMMA M, TMEM lanes)
constexpr int BNH = 64;            // output units per CTA (UMMA N)
constexpr int BN = 2 * BNH;        // output units of the GEMM
constexpr int KC = 32;             // K per chunk
constexpr int MAXCH = 12;          // K <= 384
constexpr int B_LBO = 1024, B_SBO = 128;     // weight operand [64 x 32]: 8 core-matrix columns of 8 row groups
constexpr int B_OPER = 8 * B_LBO;
constexpr int B_CHUNK = 2 * B_OPER;          // hi | lo
constexpr int NSLOT = 2;            // round slots: a round = two chunks (hops between the roles cost more than the work of one chunk)
constexpr int W_DRAIN = 8, W_ISSUE = 16;
constexpr int NT = 20 * 32;
constexpr int RAW_OFF = MAXCH * B_CHUNK;     // raw ring: 2 stages of one chunk, [128 rows x 128 B], XOR-swizzled 16-byte pieces
constexpr int RAW_BYTES = BM * KC * 4;
constexpr int BAR_OFF = RAW_OFF + 2 * RAW_BYTES;
constexpr int SMEM_BYTES = BAR_OFF + 256;
constexpr int TM_A = 0, TM_D = NSLOT * 128;   // per round slot: 2 x (32 hi | 32 lo) columns of A, 2 x 64 accumulator columns
constexpr int TMEM_COLS = 512;
// launch allocation 640 x 96 = 61440 >= 256 x 96 + 256 x 120 + 128 x 40 = 60416
constexpr int REGS_LOAD = 96, REGS_DRAIN = 120, REGS_ISSUE = 40;
constexpr int PF_AHEAD = 8;                  // L2 prefetch distance of the copy warps, in chunks

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p) : "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, float v0, float v1, float v2, float v3, float v4, float v5, float v6, float v7) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "f"(v0), "f"(v1), "f"(v2),
               "f"(v3), "f"(v4), "f"(v5), "f"(v6), "f"(v7)
               : "memory");
}
// 16 TMEM lanes x 32 columns: register 4 v2 + 2 v1 + w of thread t <-> lane t / 4 + 8 v1, column 8 v2 + 2 (t % 4) + w
__device__ __forceinline__ void tmem_ld16x256b_x4(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// Tensor map of A as a [M, K] FP32 tensor with row pitch lda, boxes of [128 rows x 32 floats], 128-byte swizzle.
// Returns false when the driver entry point is missing or rejects the shape (the caller falls back to the streaming tile).
inline bool make_a_map(CUtensorMap* tm, const float* A, int64_t lda, int64_t M, int K) {
  typedef CUresult (*Encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Encode enc = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
    return (Encode)fn;
  }();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
  const cuuint64_t strides[1] = {(cuuint64_t)lda * 4};
  const cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)BM};
  const cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(A), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// grid = 2 * npairs (CTA b: pair b / 2, output units 64 (b % 2) ..).  K % 64 == 0, 64 <= K <= 384, lda / ldb / ldc % 4 == 0.
template <bool ACCUM>
__global__ void __launch_bounds__(NT, 1) dgrad_res_kernel(const __grid_constant__ CUtensorMap tmA, const float* __restrict__ Bt,
                                                          int64_t ldb, int64_t hl_stride, float* __restrict__ C, int64_t ldc,
                                                          int64_t M, int K) {
  extern __shared__ __align__(1024) char smem[];   // (the 128-byte swizzle of the raw ring works on absolute address bits)
  const uint32_t bar0 = tc::smem_u32(smem + BAR_OFF);
  const uint32_t A_FULL = bar0, A_EMPTY = bar0 + 32, D_FULL = bar0 + 64, D_EMPTY = bar0 + 96, RAW_FULL = bar0 + 128,
                 RAW_EMPTY = bar0 + 144;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 160);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = blockIdx.x & 1;
  const int64_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int nch = K / KC;
  const int64_t ntiles = (M + BM - 1) / BM;
  const uint32_t my_tiles = (pair < ntiles) ? (uint32_t)((ntiles - pair + npairs - 1) / npairs) : 0u;
  const uint32_t G = my_tiles * (uint32_t)nch;   // chunks this CTA walks (32-bit: the host launches slabs of < 2^31 rows)
  const uint32_t NR = G >> 1;                    // rounds of two chunks (nch is even)

  if (warp == W_ISSUE) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tc::smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      tc::mbar_init(A_FULL + 8 * s, 256);    // every transformer thread (its tcgen05.st has completed)
      tc::mbar_init(A_EMPTY + 8 * s, 256);   // every drain thread, once it has seen the round's MMAs complete (D_FULL)
      tc::mbar_init(D_FULL + 8 * s, 1);      // tcgen05.commit
      tc::mbar_init(D_EMPTY + 8 * s, 256);   // every drain thread
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(RAW_FULL + 8 * s, 1);     // expect_tx arrive + the bytes of the tensor copy
      tc::mbar_init(RAW_EMPTY + 8 * s, 128);  // every thread of the stage's transformer warpgroup
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // resident weights of this half: Bt[64 half + n, k] -> chunk k / 32, hi | lo, core-matrix column (k % 32) / 4, row n
  {
    const int kq = K >> 2;   // 16-byte pieces per weight row
    for (int idx = tid; idx < BNH * kq; idx += NT) {
      const int n = idx / kq, q = idx - n * kq;
      const float* src = Bt + hl_stride + (int64_t)(half * BNH + n) * ldb + q * 4;
      const float4 h = __ldg(reinterpret_cast<const float4*>(src));
      const float4 l = __ldg(reinterpret_cast<const float4*>(src + hl_stride));
      char* dst = smem + (q >> 3) * B_CHUNK + (q & 7) * B_LBO + (n >> 3) * B_SBO + (n & 7) * 16;
      *reinterpret_cast<float4*>(dst) = h;
      *reinterpret_cast<float4*>(dst + B_OPER) = l;
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy stores -> UMMA
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp < W_DRAIN) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(REGS_LOAD));
    // ================================ transformers: raw ring -> hi / lo -> TMEM =====================================
    // thread = row of the tile = TMEM lane (tcgen05.st 32x32b); warpgroup w takes chunk 2 R + w of round R, which the
    // copy thread put into ring stage w, and writes half w of the round's A slot.
    const int wg = warp >> 2, quarter = warp & 3;
    const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16) + TM_A + wg * 64;
    const char* prow = smem + RAW_OFF + wg * RAW_BYTES + (quarter * 32 + lane) * 128;
    const int sw = (int)((tc::smem_u32(prow) >> 7) & 7);   // = lane & 7 for a 1024-byte aligned ring
    DG_DECL;
    DG_T(tl0);
#pragma unroll 1
    for (uint32_t R = 0; R < NR; ++R) {
      const int rs = (int)(R & 1);
      const uint32_t ur = R >> 1;
      DG_T(t0);
      tc::mbar_wait(RAW_FULL + 8 * wg, R & 1);
      DG_ADD(0, t0);
      DG_T(t1);
      float4 x[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) x[q] = *reinterpret_cast<const float4*>(prow + ((q ^ sw) << 4));
      float4 h[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        h[q].x = tc::tf32_hi(x[q].x); h[q].y = tc::tf32_hi(x[q].y); h[q].z = tc::tf32_hi(x[q].z); h[q].w = tc::tf32_hi(x[q].w);
      }
      // Hand the stage back -- but only once the eight loads have been PERFORMED.  mbarrier.arrive alone is not enough:
      // ptxas places SYNCS.ARRIVE right behind the LDS instructions (nothing depends on their results yet), the arrive
      // overtakes loads still queued in the shared-memory pipe, the copy thread sees the stage free and the next tensor
      // copy lands on top of the last piece before it has been read (measured: exactly that piece carried the data of
      // chunk g + 2 in ~1000 of 4.8 M results, only in the MMA-bound steady state).  The fence orders the loads first.
#ifndef DGMK_DG_PRED_RELEASE
      asm volatile("fence.acq_rel.cta;\n" ::: "memory");
      lg::mbar_arrive(RAW_EMPTY + 8 * wg);
#else
      {   // alternative, measured equal (1.105 against 1.095 ms): exactly one of two predicated arrives executes, and the
          // predicate depends on a word of every load, so neither can issue before the loads have written their registers
          // (ptxas keeps both; a data dependence it can fold -- and t, x, 0 -- it removes)
        const uint32_t dep = __float_as_uint(x[0].x) ^ __float_as_uint(x[1].y) ^ __float_as_uint(x[2].z) ^ __float_as_uint(x[3].w) ^
                             __float_as_uint(x[4].x) ^ __float_as_uint(x[5].y) ^ __float_as_uint(x[6].z) ^ __float_as_uint(x[7].w);
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0x5bd1e995;\n\t"
            "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n\t"
            "@!p mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}\n" ::"r"(RAW_EMPTY + 8 * wg),
            "r"(dep)
            : "memory");
      }
#endif
      DG_ADD(1, t1);
      DG_T(t2);
      tc::mbar_wait(A_EMPTY + 8 * rs, (ur & 1) ^ 1);   // the MMAs of round R - 2 have read this slot
      DG_ADD(2, t2);
      DG_T(t3);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint32_t ta = tlane + rs * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 xa = x[2 * q], xb = x[2 * q + 1], ha = h[2 * q], hb = h[2 * q + 1];
        tmem_st8(ta + q * 8, ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w);
        tmem_st8(ta + 32 + q * 8, xa.x - ha.x, xa.y - ha.y, xa.z - ha.z, xa.w - ha.w, xb.x - hb.x, xb.y - hb.y, xb.z - hb.z, xb.w - hb.w);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      lg::mbar_arrive(A_FULL + 8 * rs);
      DG_ADD(3, t3);
    }
    DG_ADD(4, tl0);
    if (warp == 0) { DG_OUT(0); }
  } else if (warp < W_ISSUE) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(REGS_DRAIN));
    // ================================ drain: chunk results -> RN registers -> C =====================================
    // tcgen05.ld 16x256b.x4 (two per chunk: lanes 0-15 / 16-31 of the warp's quarter, 32 columns): register
    // 4 v2 + 2 v1 + w of thread t = lane t / 4 + 8 v1, column 8 v2 + 2 (t % 4) + w -- four threads hold 32 contiguous
    // bytes of a row of C, so the read-modify-write of the tile needs no transposition: 8-byte accesses, whole sectors.
    const int quarter = warp & 3, part = (warp - W_DRAIN) >> 2;
    const uint32_t tbase = tmem + ((uint32_t)(quarter * 32) << 16) + TM_D + part * 32;
    const int col0 = half * BNH + part * 32 + 2 * (lane & 3);
    const int rpt = nch >> 1;   // rounds per tile
    float acc[32];
    int rt = 0;                 // round of the tile
    uint32_t ti = 0;
    DG_DECL;
    DG_T(td0);
#pragma unroll 1
    for (uint32_t R = 0; R < NR; ++R) {
      const int rs = (int)(R & 1);
      const uint32_t ur = R >> 1;
      const int64_t row0 = (pair + (int64_t)ti * npairs) * BM + quarter * 32 + (lane >> 2);
      float* cb = C + row0 * ldc + col0;
      if (ACCUM && rt == (rpt > 1 ? rpt - 2 : 0)) {   // this warp's part of the C tile into L2 before the read-modify-write
        const int64_t row = (pair + (int64_t)ti * npairs) * BM + quarter * 32 + lane;
        if (row < M) prefetch_l2(C + row * ldc + half * BNH + part * 32);
      }
      // last round of the tile: the thread's 16 pieces of the old C go out BEFORE the wait, so that their latency hides
      // behind the round's MMAs (issued after the tile they cost ~3000 cycles, during which the accumulator ring filled
      // up and the tensor pipe stopped)
      float2 old[16];
      const bool last = (rt == rpt - 1) && !DG_DBG(4);
      if (ACCUM && last) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {   // k = 8 hb + 2 v2 + v1
          const int dr = (k >> 3) * 16 + (k & 1) * 8;
          old[k] = (row0 + dr < M) ? *reinterpret_cast<const float2*>(cb + dr * ldc + ((k >> 1) & 3) * 8) : make_float2(0.f, 0.f);
        }
      }
      DG_T(t0);
      tc::mbar_wait(D_FULL + 8 * rs, ur & 1);
      DG_ADD(0, t0);
      DG_T(t1);
      lg::mbar_arrive(A_EMPTY + 8 * rs);   // the round's MMAs are complete: its A columns can take round R + 2
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
      for (int j = 0; j < 2; ++j) {   // the two chunk results of the round
        uint32_t v0[16], v1[16];
        tmem_ld16x256b_x4(tbase + rs * 128 + j * 64, v0);
        tmem_ld16x256b_x4(tbase + ((uint32_t)16 << 16) + rs * 128 + j * 64, v1);
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        if (j == 1) {
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
          lg::mbar_arrive(D_EMPTY + 8 * rs);   // these columns can take round R + 2
        }
        if (j == 0 && rt == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) { acc[i] = __uint_as_float(v0[i]); acc[16 + i] = __uint_as_float(v1[i]); }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) { acc[i] += __uint_as_float(v0[i]); acc[16 + i] += __uint_as_float(v1[i]); }
        }
      }
      DG_ADD(1, t1);
      if (last) {   // tile complete
        DG_T(t2);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int dr = (k >> 3) * 16 + (k & 1) * 8;
          if (row0 + dr < M) {
            float2 o = make_float2(acc[2 * k], acc[2 * k + 1]);
            if (ACCUM) { o.x += old[k].x; o.y += old[k].y; }
            *reinterpret_cast<float2*>(cb + dr * ldc + ((k >> 1) & 3) * 8) = o;
          }
        }
        DG_ADD(2, t2);
      }
      if (++rt == rpt) { rt = 0; ++ti; }
    }
    DG_ADD(3, td0);
    if (warp == W_DRAIN) { DG_OUT(8); }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(REGS_ISSUE));
    if (warp == W_ISSUE || warp == W_ISSUE + 2) {
      // ================================ MMA issuers ===================================================================
      // Two warps on alternate rounds (issuer i owns round slot i): between the last MMA of a round and the first of
      // its next one an issuer spends ~300 cycles on barriers, descriptors and the commit, longer than the tensor
      // pipe's queue lasts; with a second issuer the other round's MMAs run meanwhile (rounds are independent: own A
      // slot, own accumulator columns).
      const int iw = (warp - W_ISSUE) >> 1;
      const uint32_t sb = tc::smem_u32(smem);
      int c = (2 * iw) % nch;
      DG_DECL;
      DG_T(ti0);
#pragma unroll 1
      for (uint32_t R = (uint32_t)iw; R < NR; R += 2) {
        const uint32_t ur = R >> 1;
        DG_T(t0);
        tc::mbar_wait(D_EMPTY + 8 * iw, (ur & 1) ^ 1);   // accumulator columns drained by all 8 warps
        DG_ADD(0, t0);
        DG_T(t1);
        tc::mbar_wait(A_FULL + 8 * iw, ur & 1);
        DG_ADD(1, t1);
        DG_T(t2);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint64_t dBh = lg::make_desc(sb + (c + j) * B_CHUNK, B_LBO), dBl = lg::make_desc(sb + (c + j) * B_CHUNK + B_OPER, B_LBO);
          const uint32_t ah = tmem + TM_A + iw * 128 + j * 64, al = ah + 32;
          const uint32_t d = tmem + TM_D + iw * 128 + j * 64;
          if (lg::elect_one()) {
            if (!DG_DBG(1)) {
#pragma unroll
              for (int ks = 0; ks < KC / 8; ++ks) {   // small terms first
                const uint64_t adv = (uint64_t)((ks * 2 * B_LBO) >> 4);
                lg::mma_ts(d, al + ks * 8, dBh + adv, ks > 0 ? 1u : 0u);
                lg::mma_ts(d, ah + ks * 8, dBl + adv, 1u);
              }
#pragma unroll
              for (int ks = 0; ks < KC / 8; ++ks) {
                const uint64_t adv = (uint64_t)((ks * 2 * B_LBO) >> 4);
                lg::mma_ts(d, ah + ks * 8, dBh + adv, 1u);
              }
            }
            if (j == 1) tc::mma_commit(D_FULL + 8 * iw);   // one commit per round (a commit costs the issuer ~100 cycles)
          }
          __syncwarp();
        }
        DG_ADD(2, t2);
        c = (c + 4) % nch;
      }
      DG_ADD(3, ti0);
      if (warp == W_ISSUE) { DG_OUT(16); }
    } else if (warp == W_ISSUE + 1) {
      // ================================ copy: global -> raw ring (TMA, one tensor copy per chunk) =====================
      // A chunk is a [128 rows x 32 floats] box of the [M, K] tensor with row pitch lda.  SWIZZLE_128B puts piece p of
      // row r at slot p ^ (r & 7) of the row's 128 bytes: the transformers' row-per-thread 16-byte reads are
      // conflict-free.  Rows beyond M arrive as zeros.  (Measured on the way: the same chunk as 1024 16-byte cp.async's
      // from two warps blocks ~1500 cycles per chunk at issue, whatever level of the hierarchy serves them.)
      if (lane == 0) {
        const uint64_t tm = reinterpret_cast<uint64_t>(&tmA);
        const uint32_t dst0 = tc::smem_u32(smem + RAW_OFF);
        int c = 0;
        uint32_t ti = 0;
        int cpf = PF_AHEAD % nch;
        uint32_t tpf = PF_AHEAD / nch;
        DG_DECL;
        DG_T(tc0);
#pragma unroll 1
        for (uint32_t g = 0; g < G; ++g) {
          const int st = (int)(g & 1);
          DG_T(t0);
          tc::mbar_wait(RAW_EMPTY + 8 * st, ((g >> 1) & 1) ^ 1);   // the transformers have this stage's previous chunk in registers
          DG_ADD(0, t0);
          DG_T(t1);
          const int row0 = (int)((pair + (int64_t)ti * npairs) * BM);
          if (!DG_DBG(2)) {
            lg::mbar_expect_tx(RAW_FULL + 8 * st, RAW_BYTES);
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                             dst0 + st * RAW_BYTES),
                         "l"(tm), "r"(c * KC), "r"(row0), "r"(RAW_FULL + 8 * st)
                         : "memory");
            if (g + PF_AHEAD < G && !DG_DBG(8)) {   // a later chunk into L2
              const int rowp = (int)((pair + (int64_t)tpf * npairs) * BM);
              asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];\n" ::"l"(tm), "r"(cpf * KC), "r"(rowp) : "memory");
            }
          } else {
            lg::mbar_arrive(RAW_FULL + 8 * st);
          }
          DG_ADD(1, t1);
          if (++c == nch) { c = 0; ++ti; }
          if (++cpf == nch) { cpf = 0; ++tpf; }
        }
        DG_ADD(2, tc0);
        DG_OUT(24);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == W_ISSUE) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

}  // namespace dg
}  // namespace dgmk
