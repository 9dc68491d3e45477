// tcgen05 (5th-gen tensor core) forward / data-gradient GEMM tile, 3xTF32 split accumulation
// (sm_100a).  The weight-gradient tile lives in dgmk_gemm_tc_tn.cuh.
//
//   C[M,N] (+)= A[M,K] * Bt[N,K]^T       A, Bt row-major FP32 (both K-major), FP32 out
//
// FP32 operands are split into hi = tf32(x) and lo = x - hi (exact): A on the fly, the
// weights Bt once per step by the packing kernel (Bt_hi = Bt + hl_stride, Bt_lo = Bt +
// 2*hl_stride).  Per K chunk of 32 the tensor core forms lo*hi + hi*lo (8 MMAs, first:
// their truncation happens at 2^-11 of the magnitude) then hi*hi (4 MMAs) in TMEM starting
// from zero; lo*lo is dropped (2^-22 relative).  The tensor core adds with truncation
// (toward zero), a bias that compounds along an accumulation chain, so the chain is cut
// after every chunk: the chunk result is read back (tcgen05.ld) and added in round-to-
// nearest FP32 registers.  Measured error ~FP32 FFMA grade (profiles/r01_notes.md); the
// parity bar is 1e-5 per tensor.
//
// CTA = 128 threads = 128 tile rows (= 128 TMEM lanes), one 128x128 output tile:
//   all threads : global -> registers (one 128-byte row chunk each for A and Bt), split,
//                 store into shared memory in the UMMA canonical K-major no-swizzle layout
//                 (8x16B core matrices; LBO = K-direction stride, SBO = row-group stride)
//   thread 0    : tcgen05.mma.cta_group::1.kind::tf32, 3 MMAs per k-step of 8,
//                 tcgen05.commit -> mbarrier
//   all threads : tcgen05.ld 32x32b (thread t owns accumulator row t) -> global
// Several CTAs per SM (64 KB smem, 128 TMEM columns each) overlap staging, MMA and
// epilogue across tiles.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dgmk {
namespace tc {

constexpr int BM = 128;   // tile rows   (UMMA M)
constexpr int BN = 128;   // tile cols   (UMMA N)
constexpr int KC = 32;    // K elements staged per chunk (4 MMA k-steps of 8)
constexpr int NT = 128;
constexpr int LBO = 2048 + 16;            // bytes between core matrices adjacent in K (padded: conflict-free)
constexpr int SBO = 128;                  // bytes between 8-row groups
constexpr int OPER_BYTES = (KC / 4) * LBO;  // one [128 x 32] operand tile
constexpr int CPITCH = BN + 4;                 // floats per row of the epilogue staging tile
constexpr int SMEM_BYTES = (4 * OPER_BYTES > BM * CPITCH * 4 ? 4 * OPER_BYTES : BM * CPITCH * 4) + 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((LBO >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // layout_type = SWIZZLE_NONE (0), base_offset 0
}
// kind::tf32, D = F32, A/B = TF32, both K-major, M = 128, N = 128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
// one lane of a converged warp.  tcgen05.mma must be issued from warp-uniform control flow
// (if (warp == 0) { if (elect_one()) ... }): under a divergent `if (tid == 0)` ptxas cannot keep
// the descriptors in uniform registers and wraps EVERY MMA in an R2UR broadcast loop -- measured
// ~90 cycles per MMA on the issuing thread against 64 for the MMA itself (N = 128).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// Prefetch loads must really be ISSUED where they are written: a plain __ldg is a pure read that
// the compiler is free to sink down to its first use (it did: no load was in flight across the
// barriers and the tiles were latency-bound).  volatile asm pins the issue point.
__device__ __forceinline__ float4 ldg_f4_pinned(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
// round-to-nearest (ties away) to 10 explicit mantissa bits with two full-rate integer ops;
// cvt.rna.tf32.f32 computes the same value but runs on the slow conversion pipe (measured: the
// 32 conversions per thread per chunk cost ~1000 cycles)
__device__ __forceinline__ float tf32_hi(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
// split a float4 into hi / lo and store both 16-byte chunks
__device__ __forceinline__ void split_store(char* hi_base, char* lo_base, int off, float4 v) {
  float4 h, l;
  h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
  l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
  *reinterpret_cast<float4*>(hi_base + off) = h;
  *reinterpret_cast<float4*>(lo_base + off) = l;
}

// grid = (N / 128, ceil(M / 128)).  K % 32 == 0, N % 128 == 0, lda/ldb/ldc % 4 == 0.
template <bool ACCUM>
__global__ void __launch_bounds__(NT) gemm_nn_tc_kernel(const float* __restrict__ A, int64_t lda,
                                                        const float* __restrict__ Bt, int64_t ldb, int64_t hl_stride,
                                                        float* __restrict__ C, int64_t ldc, int64_t M, int K) {
  extern __shared__ __align__(128) char smem[];
  char* sAh = smem;
  char* sAl = smem + OPER_BYTES;
  char* sBh = smem + 2 * OPER_BYTES;
  char* sBl = smem + 3 * OPER_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SMEM_BYTES - 64);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SMEM_BYTES - 48);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const uint32_t bar_a = smem_u32(bar);

  if (warp == 0) {  // two accumulator buffers: chunk c+1 is multiplied while chunk c is drained
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "n"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(bar_a, 1);
    mbar_init(bar_a + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  // staging: float4 index i = q*128 + tid -> tile row i/8, 16-byte K chunk i%8: a warp
  // reads 4 rows x 128 contiguous bytes (coalesced); the padded LBO makes the 8 chunks
  // of a row land in distinct bank groups.
  const int srow = tid >> 3, skc = tid & 7;
  const float* abase = A + (m0 + srow) * lda + skc * 4;
  const float* bbase = Bt + hl_stride + (int64_t)(n0 + srow) * ldb + skc * 4;  // hi; lo is hl_stride further
  const int st_off = skc * LBO + (srow >> 3) * SBO + (srow & 7) * 16;   // + q * 2 * SBO per 16 rows

  float4 ra[KC / 4];
  auto load_chunk = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
      ra[q] = (m0 + srow + q * 16 < M) ? ldg_f4_pinned(abase + (int64_t)q * 16 * lda + k0)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  // weights: already split, straight copies (L2-resident, shared by every CTA)
  auto stage_b = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float* src = bbase + (int64_t)q * 16 * ldb + k0;
      float4 h = __ldg(reinterpret_cast<const float4*>(src));
      float4 l = __ldg(reinterpret_cast<const float4*>(src + hl_stride));
      *reinterpret_cast<float4*>(sBh + st_off + q * 2 * SBO) = h;
      *reinterpret_cast<float4*>(sBl + st_off + q * 2 * SBO) = l;
    }
  };
  float acc[BN];
#pragma unroll
  for (int j = 0; j < BN; ++j) acc[j] = 0.f;
  // chunk result TMEM -> registers, added in RN (cuts the tensor core's truncating chain)
  auto drain = [&](int buf) {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
    for (int cb = 0; cb < BN / 32; ++cb) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * BN + cb * 32);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[cb * 32 + j] += __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  };
  const int nchunks = K / KC;
  uint32_t phase0 = 0, phase1 = 0;   // parity of the next completion of each buffer's barrier
  load_chunk(0);
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    if (c > 0) {  // MMAs of chunk c-1 done: shared memory is free, its result sits in TMEM buffer buf^1
      if (buf) { mbar_wait(bar_a, phase0); phase0 ^= 1; } else { mbar_wait(bar_a + 8, phase1); phase1 ^= 1; }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) split_store(sAh, sAl, st_off + q * 2 * SBO, ra[q]);
    stage_b(c * KC);
    if (c + 1 < nchunks) load_chunk((c + 1) * KC);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy stores -> async proxy (UMMA)
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint64_t dAh = make_desc(smem_u32(sAh)), dAl = make_desc(smem_u32(sAl));
      const uint64_t dBh = make_desc(smem_u32(sBh)), dBl = make_desc(smem_u32(sBl));
      const uint32_t d = tmem + (uint32_t)(buf * BN);
#pragma unroll
      for (int ks = 0; ks < KC / 8; ++ks) {  // small terms first
        const uint64_t adv = (uint64_t)((ks * 2 * LBO) >> 4);  // two core matrices along K per k-step
        mma_tf32(d, dAl + adv, dBh + adv, ks > 0 ? 1u : 0u);
        mma_tf32(d, dAh + adv, dBl + adv, 1u);
      }
#pragma unroll
      for (int ks = 0; ks < KC / 8; ++ks) {
        const uint64_t adv = (uint64_t)((ks * 2 * LBO) >> 4);
        mma_tf32(d, dAh + adv, dBh + adv, 1u);
      }
      mma_commit(bar_a + 8 * buf);
      }
      __syncwarp();
    }
    if (c > 0) drain(buf ^ 1);   // overlaps the MMAs just issued
  }
  {
    const int buf = (nchunks - 1) & 1;
    if (buf) mbar_wait(bar_a + 8, phase1); else mbar_wait(bar_a, phase0);
    drain(buf);
  }

  // epilogue: thread t owns output row t.  Rows go through shared memory (the operand
  // tiles are dead now) so that each warp stores whole 512-byte rows instead of 32
  // scattered 16-byte pieces per instruction.
  {
    float* stile = reinterpret_cast<float*>(smem);
#pragma unroll
    for (int q = 0; q < BN / 4; ++q)
      *reinterpret_cast<float4*>(stile + tid * CPITCH + q * 4) = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    __syncwarp();  // each warp only re-reads the 32 rows it wrote itself
    for (int rb = 0; rb < 32; rb += 8) {   // batches of 8 rows: all loads in flight before the stores
      float4 old[8];
      if (ACCUM) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int row = warp * 32 + rb + r;
          old[r] = (m0 + row < M) ? *reinterpret_cast<const float4*>(C + (m0 + row) * ldc + n0 + lane * 4)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int row = warp * 32 + rb + r;
        if (m0 + row < M) {
          float4 o = *reinterpret_cast<const float4*>(stile + row * CPITCH + lane * 4);
          if (ACCUM) { o.x += old[r].x; o.y += old[r].y; o.z += old[r].z; o.w += old[r].w; }
          *reinterpret_cast<float4*>(C + (m0 + row) * ldc + n0 + lane * 4) = o;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(2 * BN) : "memory");
  }
}


}  // namespace tc
}  // namespace dgmk
