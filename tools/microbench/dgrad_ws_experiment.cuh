// Warp-specialised, persistent tcgen05 streaming GEMM (sm_100a), 3xTF32 split:
//
//   C[M, 128] (+)= A[M, K] * Bt[128, K]^T        K % 32 == 0 (the K = 3H data gradient of a DGM layer:
//                                                 s bar += [abar_Z | abar_G | abar_R] [W_z; W_g; W_r])
//
// Same arithmetic as dgmk_gemm_tc.cuh (per K chunk of 32: lo*hi + hi*lo then hi*hi from zero in TMEM,
// chunk results summed in round-to-nearest registers), different machine mapping.  That tile ran
// load -> split -> store -> __syncthreads -> MMA -> drain one after the other in every CTA (two CTAs
// per SM gave the only overlap): ~32 K SM-cycles per 128-row tile at K = 384 against 9.2 K of MMA
// time and 13.5 K of HBM time.  Here the roles run concurrently, decoupled by mbarriers, one CTA per
// SM walking row tiles t = blockIdx.x, + gridDim.x, ...:
//
//   warps 0-7    A stagers: coalesced 16-byte loads of the [128 x 32] row chunk (a warp instruction =
//                4 rows x 128 bytes), tf32 hi / lo split, UMMA canonical K-major tile -> 3-deep ring.
//                At the start of a tile they also pull the NEXT tile's rows of A and C into L2
//                (prefetch.global.L2), so the chunk loads see L2 latency.
//   warps 8-11   weight stagers: the chunk's pre-split hi / lo weights (L2-resident) -> 3-deep ring
//   warp 20      MMA issuer (its own warpgroup: tcgen05.mma issue blocks on the tensor pipe's queue):
//                12 MMAs (M = N = 128, K = 8, both operands from shared memory) per chunk into one of two
//                TMEM accumulator buffers; tcgen05.commit frees the ring stages and publishes the chunk
//   warps 12-19  drain: chunk result -> RN registers (64 columns per thread); at the end of a tile
//                (+)= C through a per-warp shared-memory patch, so that global accesses are 64-byte row
//                segments instead of one 16-byte piece per lane and row
// grid = (min(row tiles, SMs), N / 128).
//
// EXPERIMENT, not part of the library (tools/microbench/bench_tc.cu runs it): correct to the last bit of the
// streaming tile, but no faster -- 0.52 ms against 0.45 ms for the K = 3H data gradient of 524 288 rows.
// Both are bound by L2 bandwidth, not by their pipelines: every 128-row tile re-streams the 393 KB of
// split weights (3.1 TB/s of L2 -> SM traffic next to 2.6 TB/s for A and C).  The fix is fewer weight
// bytes per row (weights resident in the shared memory of a 4-CTA cluster that splits K and exchanges
// partial tiles through distributed shared memory), not a better pipeline -- profiles/r01_notes.md.
// Per-role cycle counters of this version (cycles per chunk, 2790 in total): A stagers 1490 in split +
// store (first use of the prefetched registers: latency-bound with one chunk in flight), weight stagers
// 1370 (same), MMA warp: issue 936, waiting 1580 for operands and accumulators; drain 709 + 968 per
// chunk for the read-modify-write of C at the end of a tile.  Variants tried on top, all slower:
// cp.async (LDGSTS) weight staging with deferred publication (0.69-0.80 ms), three register sets for
// the A stagers, C folded into the accumulators at the start of a tile.
#pragma once
#include "../../differential_equations_dnn_b200/csrc/dgmk_gemm_tc_tn.cuh"

namespace dgmk {
namespace dg {

using tc::BM; using tc::BN; using tc::KC; using tc::LBO; using tc::SBO; using tc::OPER_BYTES;

constexpr int NT = 24 * 32;
constexpr int W_B = 8, W_DRAIN = 12, W_ISSUE = 20;
constexpr int NS = 3;                                  // ring stages (A and weights advance together)
constexpr int STAGE_BYTES = 2 * OPER_BYTES;            // hi | lo
constexpr int A_OFF = 0, B_OFF = NS * STAGE_BYTES;
constexpr int PATCH_PITCH = 20;                        // floats per row of a drain warp's [32 x 16] patch
constexpr int PATCH_OFF = 2 * NS * STAGE_BYTES;
constexpr int PATCH_BYTES = 32 * PATCH_PITCH * 4;
constexpr int BAR_OFF = PATCH_OFF + 8 * PATCH_BYTES;
constexpr int SMEM_BYTES = BAR_OFF + 256;
constexpr int TMEM_COLS = 256;                         // two accumulator buffers of 128 columns
// launch allocation 768 x 80 = 61440 = 256*64 + 128*96 + 256*112 + 128*32
constexpr int REGS_A = 64, REGS_B = 96, REGS_DRAIN = 112, REGS_ISSUE = 32;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* src) {
  asm volatile("prefetch.global.L2 [%0];\n" ::"l"(src) : "memory");
}

template <bool ACCUM>
__global__ void __launch_bounds__(NT, 1) dgrad_ws_kernel(const float* __restrict__ A, int64_t lda,
                                                         const float* __restrict__ Bt, int64_t ldb, int64_t hl_stride,
                                                         float* __restrict__ C, int64_t ldc, int64_t M, int K) {
  extern __shared__ __align__(128) char smem[];
  const uint32_t bar0 = tc::smem_u32(smem + BAR_OFF);
  const uint32_t A_FULL = bar0, A_EMPTY = bar0 + 24, B_FULL = bar0 + 48, B_EMPTY = bar0 + 72, T_FULL = bar0 + 96,
                 T_EMPTY = bar0 + 112;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + BAR_OFF + 128);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * BN;
  const int64_t ntiles = (M + BM - 1) / BM;
  const int nck = K / KC;                                // chunks per tile
  const int64_t my_tiles = (ntiles > blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t G = my_tiles * nck;                      // chunks this CTA works through

  if (warp == W_DRAIN) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tc::smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      tc::mbar_init(A_FULL + 8 * s, 8);      // one arrive per A-stager warp
      tc::mbar_init(A_EMPTY + 8 * s, 1);     // tcgen05.commit
      tc::mbar_init(B_FULL + 8 * s, 4);      // one arrive per weight-stager warp
      tc::mbar_init(B_EMPTY + 8 * s, 1);     // tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(T_FULL + 8 * s, 1);      // tcgen05.commit
      tc::mbar_init(T_EMPTY + 8 * s, 256);   // every drain thread
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp < W_B) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(REGS_A));
    // ================================ A stagers ============================================
    // float4 index q*256 + tid of the chunk -> row q*32 + tid/8, 16-byte K piece tid%8
    const int srow = tid >> 3, skc = tid & 7;
    const int st_off = skc * LBO + (srow >> 3) * SBO + (srow & 7) * 16;   // + q * 4 * SBO per 32 rows
    // two register sets: the loads of chunk g+1 are in flight while chunk g is split and stored
    int stage = 0; uint32_t use = 0;
    int64_t li = 0; int lc = 0;          // (tile, chunk) of the next load
    auto load = [&](float4 (&v)[4]) {
      const int64_t m0 = ((int64_t)blockIdx.x + li * gridDim.x) * BM;
      if (lc == 0 && li + 1 < my_tiles) {   // next tile's rows of A (and of C) -> L2: row tid/2, half of its lines each
        const int64_t r = m0 + (int64_t)gridDim.x * BM + (tid >> 1);
        if (r < M) {
          const char* ap = reinterpret_cast<const char*>(A + r * lda);
          const int nl = K >> 5;   // 128-byte lines per row
          for (int l = (tid & 1); l < nl; l += 2) prefetch_l2(ap + l * 128);
          if (ACCUM) {
            const char* cp = reinterpret_cast<const char*>(C + r * ldc + n0);
            prefetch_l2(cp + (tid & 1) * 256); prefetch_l2(cp + (tid & 1) * 256 + 128);
          }
        }
      }
      const float* ap = A + (m0 + srow) * lda + skc * 4 + lc * KC;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        v[q] = (m0 + srow + q * 32 < M) ? tc::ldg_f4_pinned(ap + (int64_t)q * 32 * lda) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (++lc == nck) { lc = 0; ++li; }
    };
    auto process = [&](const float4 (&v)[4]) {
      tc::mbar_wait(A_EMPTY + 8 * stage, (use & 1) ^ 1);
      char* sh = smem + A_OFF + stage * STAGE_BYTES;
#pragma unroll
      for (int q = 0; q < 4; ++q) tc::split_store(sh, sh + OPER_BYTES, st_off + q * 4 * SBO, v[q]);
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy stores -> UMMA
      __syncwarp();
      if (lane == 0) mbar_arrive(A_FULL + 8 * stage);
      if (++stage == NS) { stage = 0; ++use; }
    };
    float4 va[4], vb[4];
    if (G > 0) load(va);
    for (int64_t g = 0; g < G; g += 2) {
      if (g + 1 < G) load(vb);
      process(va);
      if (g + 1 < G) {
        if (g + 2 < G) load(va);
        process(vb);
      }
    }
  } else if (warp < W_DRAIN) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(REGS_B));
    // ================================ weight stagers =======================================
    // pre-split weights (hi at hl_stride, lo at 2*hl_stride): straight copies, rows (tid/8) + 16 q
    const int t = tid - W_B * 32;
    const int srow = t >> 3, skc = t & 7;
    const int st_off = skc * LBO + (srow >> 3) * SBO + (srow & 7) * 16;   // + q * 2 * SBO per 16 rows
    const float* bbase = Bt + hl_stride + (int64_t)(n0 + srow) * ldb + skc * 4;
    // half chunks (rows srow + 16 q, q = 0..3 / 4..7) in two register sets: the loads of the next half are
    // in flight while this one is stored; the stage is published after its second half
    int stage = 0; uint32_t use = 0;
    int lc = 0, lh = 0;                  // (chunk within the tile, half) of the next load
    auto load = [&](float4 (&h)[4], float4 (&l)[4]) {
      const float* src = bbase + (int64_t)lh * 64 * ldb + lc * KC;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        h[q] = tc::ldg_f4_pinned(src + (int64_t)q * 16 * ldb);              // pinned: issued here
        l[q] = tc::ldg_f4_pinned(src + (int64_t)q * 16 * ldb + hl_stride);
      }
      if (++lh == 2) { lh = 0; if (++lc == nck) lc = 0; }
    };
    auto store = [&](const float4 (&h)[4], const float4 (&l)[4], int half) {
      if (half == 0) tc::mbar_wait(B_EMPTY + 8 * stage, (use & 1) ^ 1);
      char* sh = smem + B_OFF + stage * STAGE_BYTES + st_off + half * 8 * SBO;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        *reinterpret_cast<float4*>(sh + q * 2 * SBO) = h[q];
        *reinterpret_cast<float4*>(sh + OPER_BYTES + q * 2 * SBO) = l[q];
      }
      if (half == 1) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(B_FULL + 8 * stage);
        if (++stage == NS) { stage = 0; ++use; }
      }
    };
    float4 ha[4], la[4], hb[4], lb[4];
    if (G > 0) load(ha, la);
    for (int64_t g = 0; g < G; ++g) {
      load(hb, lb);                      // second half of chunk g
      store(ha, la, 0);
      if (g + 1 < G) load(ha, la);       // first half of chunk g+1
      store(hb, lb, 1);
    }
  } else if (warp < W_ISSUE) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(REGS_DRAIN));
    // ================================ drain ================================================
    const int quarter = warp & 3, half = (warp - W_DRAIN) >> 2;
    float* patch = reinterpret_cast<float*>(smem + PATCH_OFF + (warp - W_DRAIN) * PATCH_BYTES);
    float acc[tctn::HALF];
#pragma unroll
    for (int j = 0; j < tctn::HALF; ++j) acc[j] = 0.f;
    int c = 0; int64_t i = 0;
    for (int64_t g = 0; g < G; ++g) {
      const int buf = (int)(g & 1);
      tc::mbar_wait(T_FULL + 8 * buf, (uint32_t)((g >> 1) & 1));
      tctn::drain_half(tmem, buf, quarter, half, acc);   // fences inside
      mbar_arrive(T_EMPTY + 8 * buf);
      if (++c == nck) {   // tile complete: rows quarter*32 .., columns n0 + half*64 .. of C
        c = 0;
        const int64_t r0 = ((int64_t)blockIdx.x + i * gridDim.x) * BM + quarter * 32;
        ++i;
        float* cbase = C + n0 + half * tctn::HALF;
#pragma unroll
        for (int rd = 0; rd < tctn::HALF / 16; ++rd) {   // 16 columns per round through the warp's patch
          float4 old[4];
          if (ACCUM) {   // lane: row 8 k + lane/4, 16-byte piece lane%4 -- issued before the patch is written
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int64_t r = r0 + 8 * k + (lane >> 2);
              old[k] = (r < M) ? *reinterpret_cast<const float4*>(cbase + r * ldc + rd * 16 + (lane & 3) * 4)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(patch + lane * PATCH_PITCH + q * 4) =
                make_float4(acc[rd * 16 + 4 * q], acc[rd * 16 + 4 * q + 1], acc[rd * 16 + 4 * q + 2], acc[rd * 16 + 4 * q + 3]);
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int64_t r = r0 + 8 * k + (lane >> 2);
            float4 o = *reinterpret_cast<const float4*>(patch + (8 * k + (lane >> 2)) * PATCH_PITCH + (lane & 3) * 4);
            if (ACCUM) { o.x += old[k].x; o.y += old[k].y; o.z += old[k].z; o.w += old[k].w; }
            if (r < M) *reinterpret_cast<float4*>(cbase + r * ldc + rd * 16 + (lane & 3) * 4) = o;
          }
        }
#pragma unroll
        for (int j = 0; j < tctn::HALF; ++j) acc[j] = 0.f;
      }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(REGS_ISSUE));
    // ================================ MMA issuer (warp 20; 21-23 only hand their registers over) ==
    if (warp == W_ISSUE) {
      int stage = 0; uint32_t use = 0;
#pragma unroll 1
      for (int64_t g = 0; g < G; ++g) {
        const int buf = (int)(g & 1);
        const uint32_t ph = (uint32_t)((g >> 1) & 1);
        tc::mbar_wait(T_EMPTY + 8 * buf, ph ^ 1);          // accumulator buffer drained by all 8 warps
        tc::mbar_wait(A_FULL + 8 * stage, use & 1);
        tc::mbar_wait(B_FULL + 8 * stage, use & 1);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t sa = tc::smem_u32(smem + A_OFF + stage * STAGE_BYTES), sb = tc::smem_u32(smem + B_OFF + stage * STAGE_BYTES);
        const uint64_t dAh = tc::make_desc(sa), dAl = tc::make_desc(sa + OPER_BYTES);
        const uint64_t dBh = tc::make_desc(sb), dBl = tc::make_desc(sb + OPER_BYTES);
        const uint32_t d = tmem + (uint32_t)(buf * BN);
        if (tc::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < KC / 8; ++ks) {   // small terms first
            const uint64_t adv = (uint64_t)((ks * 2 * LBO) >> 4);   // two core matrices along K per k-step
            tc::mma_tf32(d, dAl + adv, dBh + adv, ks > 0 ? 1u : 0u);
            tc::mma_tf32(d, dAh + adv, dBl + adv, 1u);
          }
#pragma unroll
          for (int ks = 0; ks < KC / 8; ++ks) {
            const uint64_t adv = (uint64_t)((ks * 2 * LBO) >> 4);
            tc::mma_tf32(d, dAh + adv, dBh + adv, 1u);
          }
          tc::mma_commit(A_EMPTY + 8 * stage);
          tc::mma_commit(B_EMPTY + 8 * stage);
          tc::mma_commit(T_FULL + 8 * buf);
        }
        __syncwarp();
        if (++stage == NS) { stage = 0; ++use; }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == W_DRAIN) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

}  // namespace dg
}  // namespace dgmk
