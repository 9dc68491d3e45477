// tcgen05 (5th-gen tensor core) weight-gradient tile, 3xTF32 split accumulation (sm_100a).
// (forward / data-gradient tile: dgmk_gemm_tc.cuh; same split / drain scheme.)
//
//   gemm_tn_tc : P[z][N,Kd] = sum_{m in split z} A[m,0:N]^T S[m,0:Kd]   weight gradient
//                PE[z][4,N] = sum_m A[m,n] E[m,e]                       (input-map / bias grads)
//
// FP32 operands are split into hi = tf32(x) and lo = x - hi (exact): activations on the fly
// while they are staged, the weights Bt once per step by the packing kernel (Bt_hi = Bt +
// hl_stride, Bt_lo = Bt + 2*hl_stride).  Per K chunk of 32 the tensor core forms lo*hi + hi*lo
// (8 MMAs, first: their truncation happens at 2^-11 of the magnitude) then hi*hi (4 MMAs) in
// TMEM starting from zero; lo*lo is dropped (2^-22 relative).  The tensor core adds with
// truncation (toward zero), a bias that compounds along an accumulation chain, so the chain is
// cut after every chunk: the chunk result is read back (tcgen05.ld) and added in round-to-nearest
// FP32 registers.  Two TMEM accumulator buffers let chunk c+1 multiply while chunk c is drained.
// Measured error 1.2e-7 (the FP32 FFMA tile: 2.0e-7); the parity bar is 1e-5 per tensor.
//
// CTA = 256 threads, one 128x128 output tile = 128 TMEM lanes x 128 columns per buffer:
//   all threads : global -> registers -> tf32 split -> shared memory in the UMMA canonical
//                 layout (K-major no-swizzle for nn; MN-major SWIZZLE_128B_BASE32B for tn)
//   thread 0    : tcgen05.mma.cta_group::1.kind::tf32 x 12 per chunk, tcgen05.commit -> mbarrier
//   all threads : tcgen05.ld 32x32b; warp w reads lanes 32*(w%4).., columns 64*(w/4).. (two
//                 threads per output row, 64 accumulators each)
// Two CTAs per SM (16 warps) overlap staging, MMA and drains across tiles.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dgmk {
namespace tctn {

constexpr int BM = 128;   // tile rows   (UMMA M)
constexpr int BN = 128;   // tile cols   (UMMA N)
constexpr int KC = 32;    // K elements staged per chunk (4 MMA k-steps of 8)
constexpr int NT = 256;
constexpr int HALF = BN / 2;  // accumulator columns per thread

// ---- nn: K-major no-swizzle operand tiles -------------------------------------------------
constexpr int LBO = 2048 + 16;            // bytes between core matrices adjacent in K (padded: conflict-free)
constexpr int SBO = 128;                  // bytes between 8-row groups
constexpr int OPER_BYTES = (KC / 4) * LBO;  // one [128 x 32] operand tile
constexpr int CPITCH = BN + 4;            // floats per row of the epilogue staging tile
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
// kind::tf32, D = F32, A/B = TF32, M = 128, N = 128; bits 15/16 = A/B MN-major
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr uint32_t IDESC_MN = IDESC | (1u << 15) | (1u << 16);

template <uint32_t ID>
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(ID), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {   // see dgmk_gemm_tc.cuh
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
__device__ __forceinline__ float4 ldg_f4_pinned(const float* p) {   // see dgmk_gemm_tc.cuh
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
// TMEM setup shared by both kernels: 2 accumulator buffers of BN columns, 2 mbarriers
__device__ __forceinline__ uint32_t tc_prologue(uint64_t* bar, uint32_t* tmem_slot, int tid) {
  const uint32_t bar_a = smem_u32(bar);
  if ((tid >> 5) == 0) {
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
  return *tmem_slot;
}
__device__ __forceinline__ void tc_epilogue_free(uint32_t tmem, int tid) {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if ((tid >> 5) == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(2 * BN) : "memory");
  }
}
// chunk result TMEM -> registers, added in RN (cuts the tensor core's truncating chain):
// this thread's 64 columns of accumulator buffer `buf`
__device__ __forceinline__ void drain_half(uint32_t tmem, int buf, int quarter, int half, float (&acc)[HALF]) {
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
  for (int cb = 0; cb < HALF / 32; ++cb) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN + half * HALF + cb * 32);
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
}

// =====================================================================================
// Weight gradient.  UMMA M = 128 columns of A (output rows), UMMA N = 128 columns of S, UMMA K =
// rows; A^T E is accumulated next to it on the CUDA cores from the values being staged.  Both
// operands are MN-major straight out of the row-major activations (idesc a_major = b_major = 1).
// For 32-bit MN-major operands the only layout the tensor core accepts is SWIZZLE_128B_BASE32B
// (no-swizzle MN-major tf32 silently yields zeros -- measured): atoms of 4 rows x 128 bytes
// (32 columns), the 32-byte chunk c of row r stored at chunk c ^ (r % 4); column groups LBO
// apart, 4-row groups SBO apart.  A float4 of row m therefore lands at
//   (m/4)*SBO + (col/32)*LBO + (m%4)*128 + ((((col%32)/8) ^ (m%4))*32) + (col%8)*4
// with no transposition.  grid = (Kd / 128, N / 128, splits); N, Kd multiples of 128.
constexpr int TN_LBO = 512;                       // bytes between 32-column groups
constexpr int TN_SBO = 4 * TN_LBO;                // bytes between 4-row groups (128 columns per operand)
constexpr int TN_OPER_BYTES = (KC / 4) * TN_SBO;  // 16 KB
constexpr int TN_SMEM_BYTES = 4 * TN_OPER_BYTES + 32 + 2 * KC * 16 + 1024;

__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((TN_LBO >> 4) & 0x3FFF) << 16;  // leading: MN direction (32-column groups)
  d |= (uint64_t)((TN_SBO >> 4) & 0x3FFF) << 32;  // stride: K direction (4-row groups)
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                         // SWIZZLE_128B_BASE32B
  return d;
}
// byte offset of the 16-byte piece (row m of the chunk, columns col..col+3) inside an operand tile
__device__ __forceinline__ int tn_off(int m, int col) {
  return (m >> 2) * TN_SBO + (col >> 5) * TN_LBO + (m & 3) * 128 + ((((col & 31) >> 3) ^ (m & 3)) << 5) + ((col & 7) << 2);
}

__global__ void __launch_bounds__(NT) gemm_tn_tc_kernel(const float* __restrict__ A, int64_t lda,
                                                        const float* __restrict__ S, int64_t lds,
                                                        const float* __restrict__ E, float* __restrict__ P,
                                                        float* __restrict__ PE, int N, int Kd, int64_t M,
                                                        int64_t rows_per_split) {
  extern __shared__ char smem_raw[];
  char* smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  char* sAh = smem;
  char* sAl = smem + TN_OPER_BYTES;
  char* sBh = smem + 2 * TN_OPER_BYTES;
  char* sBl = smem + 3 * TN_OPER_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * TN_OPER_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 4 * TN_OPER_BYTES + 16);
  float4* sE = reinterpret_cast<float4*>(smem + 4 * TN_OPER_BYTES + 32);   // [2][32] rows of E

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, half = warp >> 2;
  const int i0 = blockIdx.y * BM;   // output rows  = columns of A
  const int j0 = blockIdx.x * BN;   // output cols  = columns of S
  const int64_t mb = (int64_t)blockIdx.z * rows_per_split;
  const int64_t me = (mb + rows_per_split < M) ? mb + rows_per_split : M;
  const uint32_t bar_a = smem_u32(bar);
  const bool do_e = (PE != nullptr) && (blockIdx.x == 0);
  const uint32_t tmem = tc_prologue(bar, tmem_slot, tid);

  // staging: float4 index i = q*256 + tid -> chunk row q*8 + warp, columns lane*4 .. lane*4+3
  const float* abase = A + i0 + lane * 4;
  const float* sbase = S + j0 + lane * 4;
  float4 ra[4], rs[4], re;
  auto load_chunk = [&](int64_t m0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t m = m0 + q * 8 + warp;
      const bool ok = m < me;
      ra[q] = ok ? ldg_f4_pinned(abase + m * lda) : make_float4(0.f, 0.f, 0.f, 0.f);
      rs[q] = ok ? ldg_f4_pinned(sbase + m * lds) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (do_e && tid < KC) {
      const int64_t m = m0 + tid;
      re = (m < me) ? __ldg(reinterpret_cast<const float4*>(E) + m) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  float acc[HALF];
#pragma unroll
  for (int j = 0; j < HALF; ++j) acc[j] = 0.f;
  // grad[U | b] = A^T E on the CUDA cores, from the A values this thread stages anyway:
  // eacc[x][e] for columns lane*4 + x; the eight warps hold different rows and are summed at the end
  float eacc[4][4];
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int e = 0; e < 4; ++e) eacc[x][e] = 0.f;

  const int64_t nchunks = (me - mb + KC - 1) / KC;
  uint32_t phase0 = 0, phase1 = 0;
  if (nchunks > 0) load_chunk(mb);
  for (int64_t c = 0; c < nchunks; ++c) {
    const int buf = (int)(c & 1);
    if (c > 0) {
      if (buf) { mbar_wait(bar_a, phase0); phase0 ^= 1; } else { mbar_wait(bar_a + 8, phase1); phase1 ^= 1; }
    }
    if (do_e && tid < KC) sE[buf * KC + tid] = re;   // consumed after the barrier below
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int row = q * 8 + warp;
      split_store(sAh, sAl, tn_off(row, lane * 4), ra[q]);
      split_store(sBh, sBl, tn_off(row, lane * 4), rs[q]);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint64_t dAh = make_desc_mn(smem_u32(sAh)), dAl = make_desc_mn(smem_u32(sAl));
      const uint64_t dBh = make_desc_mn(smem_u32(sBh)), dBl = make_desc_mn(smem_u32(sBl));
      const uint32_t d = tmem + (uint32_t)(buf * BN);
#pragma unroll
      for (int ks = 0; ks < KC / 8; ++ks) {  // small terms first; a k-step of 8 rows = two 4-row groups
        const uint64_t adv = (uint64_t)((ks * 2 * TN_SBO) >> 4);
        mma_tf32<IDESC_MN>(d, dAl + adv, dBh + adv, ks > 0 ? 1u : 0u);
        mma_tf32<IDESC_MN>(d, dAh + adv, dBl + adv, 1u);
      }
#pragma unroll
      for (int ks = 0; ks < KC / 8; ++ks) {
        const uint64_t adv = (uint64_t)((ks * 2 * TN_SBO) >> 4);
        mma_tf32<IDESC_MN>(d, dAh + adv, dBh + adv, 1u);
      }
      mma_commit(bar_a + 8 * buf);
      }
      __syncwarp();
    }
    if (do_e) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 ev = sE[buf * KC + q * 8 + warp];
        const float av[4] = {ra[q].x, ra[q].y, ra[q].z, ra[q].w};
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          eacc[x][0] = fmaf(av[x], ev.x, eacc[x][0]);
          eacc[x][1] = fmaf(av[x], ev.y, eacc[x][1]);
          eacc[x][2] = fmaf(av[x], ev.z, eacc[x][2]);
          eacc[x][3] = fmaf(av[x], ev.w, eacc[x][3]);
        }
      }
    }
    if (c + 1 < nchunks) load_chunk(mb + (c + 1) * KC);   // a whole iteration ahead of its use
    if (c > 0) drain_half(tmem, buf ^ 1, quarter, half, acc);
  }
  if (nchunks > 0) {
    const int buf = (int)((nchunks - 1) & 1);
    if (buf) mbar_wait(bar_a + 8, phase1); else mbar_wait(bar_a, phase0);
    drain_half(tmem, buf, quarter, half, acc);
  }
  // this thread owns output row i0 + quarter*32 + lane, columns j0 + half*64 ..
  {
    float* prow = P + (int64_t)blockIdx.z * N * Kd + (int64_t)(i0 + quarter * 32 + lane) * Kd + j0 + half * HALF;
#pragma unroll
    for (int q = 0; q < HALF / 4; ++q)
      *reinterpret_cast<float4*>(prow + q * 4) = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
  }
  if (do_e) {  // sum the eight warps' partial A^T E through shared memory (operand tiles are dead)
    float* se = reinterpret_cast<float*>(smem);   // [warp][e][128]
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
      for (int e = 0; e < 4; ++e) se[(warp * 4 + e) * BN + lane * 4 + x] = eacc[x][e];
    __syncthreads();
    if (tid < BN) {
      float* pe = PE + (int64_t)blockIdx.z * 4 * N + i0 + tid;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) s += se[(w * 4 + e) * BN + tid];
        pe[(int64_t)e * N] = s;
      }
    }
  }
  tc_epilogue_free(tmem, tid);
}

}  // namespace tctn
}  // namespace dgmk
