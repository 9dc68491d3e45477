// Weight-resident K = 3H data-gradient kernel (dgmk_dgrad_res.cuh): correctness vs an FP64 product, throughput next to
// the round-1 streaming tile, per-role cycle counters (DGMK_DG_DEBUG) and the debug switches that take one role's
// work out (bit 0: no MMAs, bit 1: no global loads, bit 2: no read-modify-write of C).
#ifndef DG_PRODUCT
#define DGMK_DG_DEBUG 1
#endif
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../differential_equations_dnn_b200/csrc/dgmk_gemm_tc.cuh"
#include "../../differential_equations_dnn_b200/csrc/dgmk_dgrad_res.cuh"
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
  for (int i = 0; i < 2; ++i) f();
  CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms / reps;
}
using namespace dgmk;

int main(int argc, char** argv) {
  CK(cudaFuncSetAttribute(dg::dgrad_res_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dg::SMEM_BYTES));
  CK(cudaFuncSetAttribute(dg::dgrad_res_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dg::SMEM_BYTES));
  CK(cudaFuncSetAttribute(tc::gemm_nn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
  CK(cudaFuncSetAttribute(tc::gemm_nn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
  const int N = 128;
  struct Case { int64_t M; int K; int64_t lda; bool accum; int pairs; };
  Case cases[] = {{128, 384, 512, false, 1}, {777, 384, 512, true, 74}, {70001, 384, 512, true, 74}, {40000, 384, 512, false, 74}, {5000, 256, 256, true, 10}, {333, 192, 512, false, 74}};
  for (auto c : cases) {
    const int64_t lda = c.lda, ldb = c.K, ldc = 128;
    std::vector<float> hA(c.M * lda), hB((size_t)N * ldb * 3), hC(c.M * ldc);
    srand(1);
    for (auto& v : hA) v = (rand() / (float)RAND_MAX - 0.5f);
    { size_t nb = (size_t)N * ldb; for (size_t i = 0; i < nb; ++i) { float v = (rand() / (float)RAND_MAX - 0.5f); union { float f; uint32_t u; } h; h.f = v; h.u = (h.u + 0x1000u) & 0xFFFFE000u; hB[i] = v; hB[nb + i] = h.f; hB[2 * nb + i] = v - h.f; } }
    for (auto& v : hC) v = (rand() / (float)RAND_MAX - 0.5f);
    float *A, *B, *C, *Cr;
    CK(cudaMalloc(&A, hA.size() * 4)); CK(cudaMalloc(&B, hB.size() * 4)); CK(cudaMalloc(&C, hC.size() * 4)); CK(cudaMalloc(&Cr, hC.size() * 4));
    CK(cudaMemcpy(A, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(B, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(C, hC.data(), hC.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(Cr, hC.data(), hC.size() * 4, cudaMemcpyHostToDevice));
    naive_nt<<<(unsigned)((c.M * N + 255) / 256), 256>>>(A, lda, B, ldb, Cr, ldc, c.M, N, c.K, c.accum);
    int64_t ntiles = (c.M + 127) / 128; int np = c.pairs < ntiles ? c.pairs : (int)ntiles;
    CUtensorMap tm;
    if (!dg::make_a_map(&tm, A, lda, c.M, c.K)) { printf("tensor map encode failed\n"); return 1; }
    if (c.accum) dg::dgrad_res_kernel<true><<<2 * np, dg::NT, dg::SMEM_BYTES>>>(tm, B, ldb, (int64_t)N * ldb, C, ldc, c.M, c.K);
    else dg::dgrad_res_kernel<false><<<2 * np, dg::NT, dg::SMEM_BYTES>>>(tm, B, ldb, (int64_t)N * ldb, C, ldc, c.M, c.K);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    std::vector<float> r1(hC.size()), r2(hC.size());
    CK(cudaMemcpy(r1.data(), C, hC.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(r2.data(), Cr, hC.size() * 4, cudaMemcpyDeviceToHost));
    printf("dgrad_res M=%ld K=%d lda=%ld accum=%d pairs=%d relerr %.3e\n", (long)c.M, c.K, (long)lda, c.accum, np, relerr(r1, r2));
    cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(Cr);
  }
  for (int64_t M : {1LL << 21, 1LL << 15}) {   // 4.3 GB of A (streams from HBM) / 64 MB (stays in L2 across the repetitions)
    const int K = 384;
    float *A, *B, *C;
    CK(cudaMalloc(&A, M * 512 * 4)); CK(cudaMalloc(&B, 3 * 128 * 384 * 4)); CK(cudaMalloc(&C, M * 128 * 4));
    CK(cudaMemset(A, 0, M * 512 * 4)); CK(cudaMemset(B, 0, 3 * 128 * 384 * 4)); CK(cudaMemset(C, 0, M * 128 * 4));
    dim3 grid(1, (unsigned)(M / tc::BM));
    float ms = time_ms([&] { tc::gemm_nn_tc_kernel<true><<<grid, tc::NT, tc::SMEM_BYTES>>>(A, 512, B, K, (int64_t)128 * 384, C, 128, M, K); }, 5);
    printf("M=%ld streaming tile (accum): %.3f ms  %.1f TFLOP/s  %.0f GB/s\n", (long)M, ms, 2.0 * M * N * K / ms * 1e-9, (double)M * (K + 2 * N) * 4 / ms * 1e-6);
    CUtensorMap tm;
    if (!dg::make_a_map(&tm, A, 512, M, K)) { printf("tensor map encode failed\n"); return 1; }
#ifdef DG_PRODUCT
    for (int dbg : {0}) {
#else
    for (int dbg : {0, 8, 1, 2, 4, 6, 7}) {
      CK(cudaMemcpyToSymbol(dg::g_dg_dbg, &dbg, 4));
#endif
      ms = time_ms([&] { dg::dgrad_res_kernel<true><<<148, dg::NT, dg::SMEM_BYTES>>>(tm, B, K, (int64_t)128 * 384, C, 128, M, K); }, 5);
      CK(cudaGetLastError());
      printf("[dbg=%d] resident (accum): %.3f ms  %.1f TFLOP/s  %.0f GB/s\n", dbg, ms, 2.0 * M * N * K / ms * 1e-9, (double)M * (K + 2 * N) * 4 / ms * 1e-6);
#ifdef DGMK_DG_PROF
      long long h[32]; CK(cudaMemcpyFromSymbol(h, dg::g_dg_prof, sizeof(h)));
      const double n = (double)((M / 128 + 73) / 74) * 12;   // chunks of CTA 0
      printf("     cycles/chunk (CTA 0): transformer(warp 0, every 2nd chunk)[wait_raw_full %.0f lds+hi %.0f wait_a_empty %.0f st %.0f total %.0f]  "
             "issuer[wait_d_empty %.0f wait_a_full %.0f issue %.0f total %.0f]  drain[wait_d_full %.0f ld+add %.0f epilogue %.0f total %.0f]  copy[wait_raw_empty %.0f issue %.0f total %.0f]\n",
             h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[16] / n, h[17] / n, h[18] / n, h[19] / n, h[8] / n, h[9] / n, h[10] / n, h[11] / n, h[24] / n, h[25] / n, h[26] / n);
#endif
    }
    cudaFree(A); cudaFree(B); cudaFree(C);
  }
  return 0;
}
