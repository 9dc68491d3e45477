// tcgen05.mma kind::tf32 issue-to-completion rate vs N and operand layout (K-major, no swizzle).
#include <cstdio>
#include <cstdlib>
#include "../../differential_equations_dnn_b200/csrc/dgmk_gemm_tc.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
using namespace dgmk::tc;
__device__ __forceinline__ uint64_t mkdesc(uint32_t saddr, int lbo, int sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_id(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128) mma_rate(long long* out, int reps, int N, int lboA, int lboB, int chain, int sameop, int ts) {
  extern __shared__ __align__(1024) char smem[];
  const int OFFB = 8 * 2064 * 4;   // B operands start here
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 200 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 200 * 1024 + 16);
  int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  uint32_t bar_a = smem_u32(bar);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) { mbar_init(bar_a, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  uint32_t tmem = *slot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        // sameop: every MMA reads the same smem; else walk 4 k-steps x 3 (hi/lo) operand copies like the real kernels
        const int ks = k & 3, term = k >> 2;
        uint32_t aoff = sameop ? 0 : ks * 2 * lboA + (term & 1) * 8 * lboA;
        uint32_t boff = sameop ? 0 : ks * 2 * lboB + (term >> 1) * 8 * lboB;
        uint64_t dA = mkdesc(smem_u32(smem) + aoff, lboA, 128), dB = mkdesc(smem_u32(smem + OFFB) + boff, lboB, 128);
        if (ts) mma_ts(tmem + 256 + (uint32_t)((chain ? 0 : (k % 4)) * 64), tmem + (term & 1) * 128 + ks * 8, dB, idesc, k >= (chain ? 1 : 4));
        else mma_id(tmem + (uint32_t)((chain ? 0 : (k % 4)) * N), dA, dB, idesc, k >= (chain ? 1 : 4));
      }
    }
    mma_commit(bar_a);
    mbar_wait(bar_a, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512) : "memory");
}
int main() {
  long long* out; CK(cudaMalloc(&out, 148 * 8));
  int smem = 200 * 1024 + 64;
  CK(cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int reps = 2000;
  for (int ts : {1, 0})
  for (int sameop : {0})
  for (int chain : {1, 0})
  for (int N : {32, 64, 128, 256}) {
    if (!chain && N == 256) continue;
    if (ts && N > 64) continue;
    for (int lb : {2}) {
      int lboA = lb == 1 ? 2064 : 2048, lboB = (lb == 0 ? N * 16 : N * 16 + 16);
      mma_rate<<<148, 128, smem>>>(out, reps, N, lboA, lboB, chain, sameop, ts);
      CK(cudaDeviceSynchronize());
      long long h[148]; CK(cudaMemcpy(h, out, 148 * 8, cudaMemcpyDeviceToHost));
      printf("ts=%d sameop=%d chain=%d N=%3d lboA=%d lboB=%d: %.1f cycles per MMA (M128 K8 tf32)\n", ts, sameop, chain, N, lboA, lboB, (double)h[0] / (reps * 12.0));
    }
  }
  return 0;
}
