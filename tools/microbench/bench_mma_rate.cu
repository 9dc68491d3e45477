// How long does one tcgen05.mma kind::tf32 M=128 N=128 K=8 take when issued back to back?
// (operands: K-major no-swizzle tiles as in dgmk_gemm_tc.cuh, and a 128B-swizzled variant)
#include <cstdio>
#include <cstdlib>
#include "../../differential_equations_dnn_b200/csrc/dgmk_gemm_tc.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
using namespace dgmk::tc;

__global__ void __launch_bounds__(128) mma_rate(long long* out, int reps, int nbuf, int swz) {
  extern __shared__ __align__(1024) char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * OPER_BYTES + 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 4 * OPER_BYTES + 1024 + 16);
  int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 4 * OPER_BYTES / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  uint32_t bar_a = smem_u32(bar);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) { mbar_init(bar_a, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  uint32_t tmem = *slot;
  if (tid == 0) {
    uint64_t dA = make_desc(smem_u32(smem)), dB = make_desc(smem_u32(smem + 2 * OPER_BYTES));
    if (swz) {  // SWIZZLE_128B K-major: 8 rows x 128 B atoms, SBO = 1024 B, LBO unused (=1)
      dA = ((uint64_t)((smem_u32(smem) >> 4) & 0x3FFF)) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      dB = ((uint64_t)((smem_u32(smem + 2 * OPER_BYTES) >> 4) & 0x3FFF)) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    }
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        uint64_t adv = swz ? (uint64_t)(((k & 3) * 32) >> 4) : (uint64_t)(((k & 3) * 2 * LBO) >> 4);
        mma_tf32(tmem + (uint32_t)((r % nbuf) * 128), dA + adv, dB + adv, k > 0);
      }
    }
    mma_commit(bar_a);
    mbar_wait(bar_a, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(256) : "memory");
}
int main() {
  long long* out; CK(cudaMalloc(&out, 148 * 8));
  int smem = 4 * OPER_BYTES + 2048;
  CK(cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int swz = 0; swz < 2; ++swz)
    for (int grid : {1, 148}) {
      int reps = 2000;
      mma_rate<<<grid, 128, smem>>>(out, reps, 2, swz);
      CK(cudaDeviceSynchronize());
      long long h[148]; CK(cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost));
      printf("swizzle=%d grid=%d: %.1f cycles per tcgen05.mma (tf32 M128 N128 K8)  => %.1f TFLOP/s tf32 per SM-set\n", swz, grid,
             (double)h[0] / (reps * 12.0), 2.0 * 128 * 128 * 8 / ((double)h[0] / (reps * 12.0)) * 1.9e9 * grid * 1e-12);
    }
  return 0;
}
