// Units-on-lanes warp-specialised tcgen05 GEMM (dgmk_lane_gemm.cuh): correctness vs FP64 naive and
// throughput at the heat/DGM shapes, next to the streaming tile of dgmk_gemm_tc.cuh.
#define DGMK_LG_DEBUG 1
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <algorithm>
#include "../../differential_equations_dnn_b200/csrc/dgmk_gemm_tc.cuh"
#include "../../differential_equations_dnn_b200/csrc/dgmk_lane_gemm.cuh"
#include "../../differential_equations_dnn_b200/csrc/dgmk_lane_epi.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void naive_nt(const float* A, int64_t lda, const float* Bt, int64_t ldb, float* C, int64_t ldc, int64_t M, int N, int K, bool accum) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  int64_t m = idx / N; int n = idx % N;
  double s = accum ? C[m * ldc + n] : 0.0;
  for (int k = 0; k < K; ++k) s += (double)A[m * lda + k] * Bt[(int64_t)n * ldb + k];
  C[m * ldc + n] = (float)s;
}
static double relerr(const std::vector<float>& a, const std::vector<float>& b) {
  double num = 0, den = 0;
  for (size_t i = 0; i < a.size(); ++i) { double d = (double)a[i] - b[i]; num += d * d; den += (double)b[i] * b[i]; }
  return sqrt(num / (den + 1e-300));
}
template <typename F> float time_ms(F f, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) f();
  CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms / reps;
}
using namespace dgmk;
static XSrc xsrc1(const float* p, int64_t rows, int d) {
  XSrc x; x.p[0] = p; x.p[1] = x.p[2] = nullptr; x.block_rows = rows > 0 ? rows : 1; x.block_stride = 0; x.nptr = 1; x.d = d;
  return x;
}
struct NoStoreEpi {
  float* C; int64_t ldc;
  struct Const { int col; };
  struct Tile {};
  struct State {};
  struct Pre {};
  __device__ __forceinline__ void finish(const State&, const Const&, int) const {}
  __device__ __forceinline__ Const init(int gate, int j) const { Const k; k.col = gate * 128 + j; return k; }
  __device__ __forceinline__ void tile(Tile&, const Const&, int64_t, int64_t, int) const {}
  template <bool FULL>
  __device__ __forceinline__ void prefetch(Pre&, const Tile&, const Const&, int64_t, int, int64_t) const {}
  template <bool FULL>
  __device__ __forceinline__ void apply(const Pre&, State&, const Tile&, int, const Const& k, int64_t row0, int64_t M, const float (&a)[8]) const {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += a[q];
    if (s == 12345.678f) C[row0 * ldc + k.col] = s;
  }
};
int main() {
  CK(cudaFuncSetAttribute(lg::lane_gemm_kernel<NoStoreEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize, lg::SMEM_BYTES));
  CK(cudaFuncSetAttribute(tc::gemm_nn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
  CK(cudaFuncSetAttribute(lg::lane_gemm_kernel<lg::StoreEpi<false>>, cudaFuncAttributeMaxDynamicSharedMemorySize, lg::SMEM_BYTES));
  CK(cudaFuncSetAttribute(lg::lane_gemm_kernel<lg::StoreEpi<true>>, cudaFuncAttributeMaxDynamicSharedMemorySize, lg::SMEM_BYTES));
  struct Case { int64_t M; int N; int64_t lda; bool accum; int grid; };
  Case cases[] = {{64, 128, 128, false, 1}, {1000, 384, 128, false, 147}, {4133, 384, 512, false, 6}, {777, 128, 160, true, 148},
                  {70000, 128, 128, false, 148}};
  const int K = 128;
  for (auto c : cases) {
    int64_t lda = c.lda, ldb = K, ldc = c.N + 64;
    std::vector<float> hA(c.M * lda), hB((size_t)c.N * ldb * 3), hC(c.M * ldc);
    srand(1);
    for (auto& v : hA) v = (rand() / (float)RAND_MAX - 0.5f);
    { size_t nb = (size_t)c.N * ldb; for (size_t i = 0; i < nb; ++i) { float v = (rand() / (float)RAND_MAX - 0.5f); union { float f; uint32_t u; } h; h.f = v; h.u = (h.u + 0x1000u) & 0xFFFFE000u; hB[i] = v; hB[nb + i] = h.f; hB[2 * nb + i] = v - h.f; } }
    for (auto& v : hC) v = (rand() / (float)RAND_MAX - 0.5f);
    float *A, *B, *C, *Cr;
    CK(cudaMalloc(&A, hA.size() * 4)); CK(cudaMalloc(&B, hB.size() * 4));
    CK(cudaMalloc(&C, hC.size() * 4)); CK(cudaMalloc(&Cr, hC.size() * 4));
    CK(cudaMemcpy(A, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(B, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(C, hC.data(), hC.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(Cr, hC.data(), hC.size() * 4, cudaMemcpyHostToDevice));
    naive_nt<<<(unsigned)((c.M * c.N + 255) / 256), 256>>>(A, lda, B, ldb, Cr, ldc, c.M, c.N, K, c.accum);
    int ng = c.N / 128;
    int grid = c.grid / ng * ng; if (grid < ng) grid = ng;
    if (c.accum) { lg::StoreEpi<true> e{C, ldc}; lg::lane_gemm_kernel<<<grid, lg::NT, lg::SMEM_BYTES>>>(A, lda, B, ldb, (int64_t)c.N * ldb, c.M, ng, grid / 3, grid / 3, e); }
    else { lg::StoreEpi<false> e{C, ldc}; lg::lane_gemm_kernel<<<grid, lg::NT, lg::SMEM_BYTES>>>(A, lda, B, ldb, (int64_t)c.N * ldb, c.M, ng, grid / 3, grid / 3, e); }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> r1(hC.size()), r2(hC.size());
    CK(cudaMemcpy(r1.data(), C, hC.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(r2.data(), Cr, hC.size() * 4, cudaMemcpyDeviceToHost));
    printf("lane_gemm M=%ld N=%d lda=%ld accum=%d grid=%d relerr %.3e   C[0..3]= %g %g %g %g  ref %g %g %g %g\n", (long)c.M, c.N, (long)lda,
           c.accum, grid, relerr(r1, r2), r1[0], r1[1], r1[2], r1[3], r2[0], r2[1], r2[2], r2[3]);
    cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(Cr);
  }
  {
    int64_t M = 4LL << 17;
    float *A, *B, *C;
    CK(cudaMalloc(&A, M * 512 * 4)); CK(cudaMalloc(&B, 3 * 512 * 512 * 4)); CK(cudaMalloc(&C, M * 512 * 4));
    CK(cudaMemset(A, 0, M * 512 * 4)); CK(cudaMemset(B, 0, 3 * 512 * 512 * 4));
    struct T { const char* name; int N; int64_t lda; int64_t ldc; } ts[] = {
        {"fwd ZGR  [M,128]x[128,384] lda=128 ldc=512", 384, 128, 512}, {"fwd H    [M,128]x[128,128] lda=128 ldc=512", 128, 128, 512},
        {"fwd ZGR  [M,128]x[128,384] lda=128 ldc=384", 384, 128, 384}, {"fwd H    [M,128]x[128,128] lda=128 ldc=128", 128, 128, 128}};
    for (int dbg : {0})
    for (auto t : ts) {
      
      { int d = dbg & 3; CK(cudaMemcpyToSymbol(lg::g_lg_dbg, &d, 4)); }
      const bool nostore = dbg >= 4;
      printf("[dbg=%d%s] ", dbg & 3, nostore ? " nostore" : "");
      dim3 grid(t.N / tc::BN, (unsigned)(M / tc::BM));
      float ms = time_ms([&] { tc::gemm_nn_tc_kernel<false><<<grid, tc::NT, tc::SMEM_BYTES>>>(A, t.lda, B, K, (int64_t)512 * 512, C, t.ldc, M, K); }, 10);
      printf("%s: streaming tile %.3f ms  %.2f TFLOP/s (fp32-equivalent)", t.name, ms, 2.0 * M * t.N * K / ms * 1e-9);
      int ng = t.N / 128, g = 148 / ng * ng;
      lg::StoreEpi<false> e{C, t.ldc};
      NoStoreEpi e2{C, 512};
      if (nostore) ms = time_ms([&] { lg::lane_gemm_kernel<<<g, lg::NT, lg::SMEM_BYTES>>>(A, t.lda, B, K, (int64_t)512 * 512, M, ng, g / 3, g / 3, e2); }, 10);
      else
      ms = time_ms([&] { lg::lane_gemm_kernel<<<g, lg::NT, lg::SMEM_BYTES>>>(A, t.lda, B, K, (int64_t)512 * 512, M, ng, g / 3, g / 3, e); }, 10);
      double bytes = (double)M * (128 + t.N) * 4;
      printf("   lane_gemm %.3f ms  %.2f TFLOP/s  %.0f GB/s (algorithmic)\n", ms, 2.0 * M * t.N * K / ms * 1e-9, bytes / ms * 1e-6);
      { long long h[32]; CK(cudaMemcpyFromSymbol(h, lg::g_lg_prof, sizeof(h)));
        double n = (double)h[2];
        printf("      cycles/tile (CTA 0, %d tiles): copy[wait_raw_empty %.0f total %.0f]  mma[wait_tcempty %.0f wait_opfull %.0f issue %.0f]  transform[wait_rawfull %.0f wait_opempty %.0f work %.0f]  epi[wait_tcfull %.0f drain %.0f epi %.0f]\n",
               (int)n, h[0] / n, h[1] / n, h[8] / n, h[9] / n, h[10] / n, h[16] / n, h[17] / n, h[18] / n, h[24] / n, h[25] / n, h[26] / n); }
    }
    CK(cudaGetLastError());
    // ---- the fused DGM stages at the heat shapes (synthetic buffers) ----
    {
      int d0 = 0; CK(cudaMemcpyToSymbol(lg::g_lg_dbg, &d0, 4));
      float *X, *UB, *S2, *SR, *SN;
      CK(cudaMalloc(&X, M * 2 * 4)); CK(cudaMalloc(&UB, 512 * 16)); CK(cudaMalloc(&S2, M * 128 * 4)); CK(cudaMalloc(&SR, M * 128 * 4)); CK(cudaMalloc(&SN, M * 128 * 4));
      CK(cudaMemset(X, 0, M * 2 * 4)); CK(cudaMemset(UB, 0, 512 * 16)); CK(cudaMemset(S2, 0, M * 128 * 4)); CK(cudaMemset(SR, 0, M * 128 * 4));
      const int c0g = 47;   // 47 | 47 | 54 CTAs for Z | G | R
      auto prof = [&](const char* name, float ms, double bytes) {
        long long h[32]; CK(cudaMemcpyFromSymbol(h, lg::g_lg_prof, sizeof(h))); double n = (double)h[2];
        printf("%s: %.3f ms  %.0f GB/s\n      cycles/tile (CTA 0, %d tiles): copy[wait_raw_empty %.0f total %.0f]  mma[wait_tcempty %.0f wait_opfull %.0f issue %.0f]  transform[wait_rawfull %.0f wait_opempty %.0f work %.0f]  epi[wait_tcfull %.0f drain %.0f epi %.0f]\n",
               name, ms, bytes / ms * 1e-6, (int)n, h[0] / n, h[1] / n, h[8] / n, h[9] / n, h[10] / n, h[16] / n, h[17] / n, h[18] / n, h[24] / n, h[25] / n, h[26] / n);
      };
      {
        using E1 = lg::DgmFwd1Epi<CsHeat, ACT_TANH>;
        CK(cudaFuncSetAttribute(lg::lane_gemm_kernel<E1>, cudaFuncAttributeMaxDynamicSharedMemorySize, lg::SMEM_BYTES));
        E1 e; e.xs = xsrc1(X, M / 4, 2); e.A4 = C; e.ub = (const F4*)UB; e.S = S2; e.SR = SR;
        float ms = time_ms([&] { lg::lane_gemm_kernel<<<148, lg::NT, lg::SMEM_BYTES>>>(S2, 128, B, K, (int64_t)512 * 512, M, 3, c0g, c0g, e); }, 10);
        prof("fused Fwd1 heat (Z|G|R + act + s*R)", ms, (double)M * 512 * 5);
        int c2 = 100; CK(cudaMemcpyToSymbol(lg::g_lg_prof_cta, &c2, 4));
        ms = time_ms([&] { lg::lane_gemm_kernel<<<148, lg::NT, lg::SMEM_BYTES>>>(S2, 128, B, K, (int64_t)512 * 512, M, 3, c0g, c0g, e); }, 10);
        prof("   same, counters of CTA 2 (R gate)", ms, (double)M * 512 * 5);
      }
      {
        using E1 = lg::DgmFwd1Epi<CsV, ACT_TANH>;
        CK(cudaFuncSetAttribute(lg::lane_gemm_kernel<E1>, cudaFuncAttributeMaxDynamicSharedMemorySize, lg::SMEM_BYTES));
        E1 e; e.xs = xsrc1(X, M, 2); e.A4 = C; e.ub = (const F4*)UB; e.S = S2; e.SR = SR;
        float ms = time_ms([&] { lg::lane_gemm_kernel<<<148, lg::NT, lg::SMEM_BYTES>>>(S2, 128, B, K, (int64_t)512 * 512, M, 3, c0g, c0g, e); }, 10);
        prof("   value rows, CTA 2 (R gate)", ms, (double)M * 512 * 5);
        int c0 = 0; CK(cudaMemcpyToSymbol(lg::g_lg_prof_cta, &c0, 4));
        ms = time_ms([&] { lg::lane_gemm_kernel<<<148, lg::NT, lg::SMEM_BYTES>>>(S2, 128, B, K, (int64_t)512 * 512, M, 3, c0g, c0g, e); }, 10);
        prof("fused Fwd1 value rows (CTA 0, Z gate)", ms, (double)M * 512 * 5);
      }
      {
        using E2 = lg::DgmFwd2Epi<CsHeat, ACT_TANH>;
        CK(cudaFuncSetAttribute(lg::lane_gemm_kernel<E2>, cudaFuncAttributeMaxDynamicSharedMemorySize, lg::SMEM_BYTES));
        E2 e; e.xs = xsrc1(X, M / 4, 2); e.A4 = C; e.ub = (const F4*)UB; e.S = S2; e.Sn = SN;
        float ms = time_ms([&] { lg::lane_gemm_kernel<<<148, lg::NT, lg::SMEM_BYTES>>>(SR, 128, B, K, (int64_t)512 * 512, M, 1, 0, 0, e); }, 10);
        prof("fused Fwd2 heat (H + act + state update)", ms, (double)M * 512 * 6);
      }
    }
  }
  printf("done\n");
  return 0;
}
