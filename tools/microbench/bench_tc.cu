// tcgen05 3xTF32 GEMM tile: correctness vs FP64 naive + throughput at the heat/DGM shapes.
#define DGMK_WG_DEBUG 1
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../differential_equations_dnn_b200/csrc/dgmk_gemm_tc.cuh"
#include "../../differential_equations_dnn_b200/csrc/dgmk_gemm_tc_tn.cuh"
#include "../../differential_equations_dnn_b200/csrc/dgmk_wgrad_ws.cuh"
#include "dgrad_ws_experiment.cuh"
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
int main() {
  using namespace dgmk::tc;
  CK(cudaFuncSetAttribute(gemm_nn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  CK(cudaFuncSetAttribute(gemm_nn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  struct Case { int64_t M; int N, K; bool accum; };
  Case cases[] = {{128, 128, 32, false}, {128, 128, 128, false}, {1000, 384, 128, false}, {777, 128, 384, true}, {1500, 128, 384, true}, {1000, 128, 128, false}};
  for (auto c : cases) {
    int64_t lda = c.K + 32, ldb = c.K, ldc = c.N + 64;
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
    naive_nt<<<(unsigned)((c.M * c.N + 255) / 256), 256>>>(A, lda, B, ldb, Cr, ldc, c.M, c.N, c.K, c.accum);
    dim3 grid(c.N / BN, (unsigned)((c.M + BM - 1) / BM));
    if (c.accum) gemm_nn_tc_kernel<true><<<grid, NT, SMEM_BYTES>>>(A, lda, B, ldb, (int64_t)c.N * ldb, C, ldc, c.M, c.K);
    else gemm_nn_tc_kernel<false><<<grid, NT, SMEM_BYTES>>>(A, lda, B, ldb, (int64_t)c.N * ldb, C, ldc, c.M, c.K);
    CK(cudaDeviceSynchronize());
    std::vector<float> r1(hC.size()), r2(hC.size());
    CK(cudaMemcpy(r1.data(), C, hC.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(r2.data(), Cr, hC.size() * 4, cudaMemcpyDeviceToHost));
    printf("gemm_nn_tc M=%ld N=%d K=%d accum=%d relerr %.3e   C[0..3]= %g %g %g %g  ref %g %g %g %g\n", (long)c.M, c.N, c.K, c.accum,
           relerr(r1, r2), r1[0], r1[1], r1[2], r1[3], r2[0], r2[1], r2[2], r2[3]);
    // the warp-specialised persistent variant on the same case (few CTAs: every CTA walks several tiles)
    if (c.N == 128) {
      CK(cudaFuncSetAttribute(dgmk::dg::dgrad_ws_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dgmk::dg::SMEM_BYTES));
      CK(cudaFuncSetAttribute(dgmk::dg::dgrad_ws_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dgmk::dg::SMEM_BYTES));
      CK(cudaMemcpy(C, hC.data(), hC.size() * 4, cudaMemcpyHostToDevice));
      dim3 g2(3, c.N / BN);
      if (c.accum) dgmk::dg::dgrad_ws_kernel<true><<<g2, dgmk::dg::NT, dgmk::dg::SMEM_BYTES>>>(A, lda, B, ldb, (int64_t)c.N * ldb, C, ldc, c.M, c.K);
      else dgmk::dg::dgrad_ws_kernel<false><<<g2, dgmk::dg::NT, dgmk::dg::SMEM_BYTES>>>(A, lda, B, ldb, (int64_t)c.N * ldb, C, ldc, c.M, c.K);
      CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(r1.data(), C, hC.size() * 4, cudaMemcpyDeviceToHost));
      printf("dgrad_ws   M=%ld N=%d K=%d accum=%d relerr %.3e\n", (long)c.M, c.N, c.K, c.accum, relerr(r1, r2));
    }
    cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(Cr);
  }
  {
    int64_t M = 4LL << 17;
    float *A, *B, *C;
    CK(cudaMalloc(&A, M * 512 * 4)); CK(cudaMalloc(&B, 3 * 512 * 512 * 4)); CK(cudaMalloc(&C, M * 512 * 4));
    CK(cudaMemset(A, 0, M * 512 * 4)); CK(cudaMemset(B, 0, 3 * 512 * 512 * 4));
    struct T { const char* name; int N, K; bool acc; } ts[] = {
        {"fwd ZGR  [M,128]x[128,384]", 384, 128, false}, {"fwd H    [M,128]x[128,128]", 128, 128, false},
        {"dgrad ZGR[M,384]x[384,128]", 128, 384, true}};
    for (auto t : ts) {
      dim3 grid(t.N / BN, (unsigned)(M / BM));
      float ms = time_ms([&] {
        if (t.acc) gemm_nn_tc_kernel<true><<<grid, NT, SMEM_BYTES>>>(A, 512, B, t.K, (int64_t)512 * 512, C, 512, M, t.K);
        else gemm_nn_tc_kernel<false><<<grid, NT, SMEM_BYTES>>>(A, 512, B, t.K, (int64_t)512 * 512, C, 512, M, t.K);
      }, 10);
      printf("%s: %.3f ms  %.2f TFLOP/s (fp32-equivalent)\n", t.name, ms, 2.0 * M * t.N * t.K / ms * 1e-9);
    }
    {
      float ms = time_ms([&] { dgmk::dg::dgrad_ws_kernel<true><<<dim3(148, 1), dgmk::dg::NT, dgmk::dg::SMEM_BYTES>>>(A, 512, B, 384, (int64_t)512 * 512, C, 512, M, 384); }, 10);
      printf("dgrad ZGR[M,384]x[384,128] warp-specialised persistent: %.3f ms  %.2f TFLOP/s (fp32-equivalent)\n", ms, 2.0 * M * 128 * 384 / ms * 1e-9);
    }
    {
      CK(cudaFuncSetAttribute(dgmk::tctn::gemm_tn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dgmk::tctn::TN_SMEM_BYTES));
      int64_t rps = 4096; int splits = (int)(M / rps);
      float *P, *PE, *E;
      CK(cudaMalloc(&P, (size_t)splits * 384 * 128 * 4)); CK(cudaMalloc(&PE, (size_t)splits * 4 * 384 * 4)); CK(cudaMalloc(&E, M * 16));
      CK(cudaMemset(E, 0, M * 16));
      dim3 g2(1, 3, splits);
      float ms = time_ms([&] { dgmk::tctn::gemm_tn_tc_kernel<<<g2, dgmk::tctn::NT, dgmk::tctn::TN_SMEM_BYTES>>>(C, 512, A, 512, E, P, PE, 384, 128, M, rps); }, 10);
      printf("wgrad ZGR [M,384]^T x [M,128] (tc): %.3f ms  %.2f TFLOP/s (fp32-equivalent)\n", ms, 2.0 * M * 384 * 128 / ms * 1e-9);
      CK(cudaFuncSetAttribute(dgmk::wg::wgrad_ws_kernel<512, 512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dgmk::wg::SMEM_BYTES));
      CK(cudaFuncSetAttribute(dgmk::wg::wgrad_ws_kernel<512, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dgmk::wg::SMEM_BYTES_SEP));
      for (int sp : {49, 98, 128}) {
        int64_t rps2 = ((M + sp - 1) / sp + 31) / 32 * 32; int splits2 = (int)((M + rps2 - 1) / rps2);
        if (splits2 > splits) continue;
        dim3 g3(1, 3, splits2);
        ms = time_ms([&] { dgmk::wg::wgrad_ws_kernel<512, 512, false><<<g3, dgmk::wg::NT, dgmk::wg::SMEM_BYTES>>>(C, A, E, P, PE, 384, 128, M, rps2, 1); }, 10);
        printf("wgrad ZGR warp-specialised, %d splits: %.3f ms  %.2f TFLOP/s\n", splits2, ms, 2.0 * M * 384 * 128 / ms * 1e-9);
        { long long h[32]; CK(cudaMemcpyFromSymbol(h, dgmk::wg::g_wg_prof, sizeof(h))); double n = (double)h[30];
          printf("   cycles/chunk (%d chunks): loader[issue_loads %.0f wait_aempty %.0f split+st %.0f E %.0f] stager[loads|wait_raw %.0f wait_bempty %.0f split+sts %.0f] mma[wait_tempty %.0f wait_afull %.0f wait_bfull %.0f issue %.0f] drain[wait_tfull %.0f drain %.0f]\n",
                 (int)n, h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[8] / n, h[9] / n, h[10] / n, h[26] / n, h[27] / n, h[28] / n, h[29] / n, h[24] / n, h[25] / n); }
        // the same launch without the A^T E side product (the fused path forms grad[U | b] elsewhere)
        ms = time_ms([&] { dgmk::wg::wgrad_ws_kernel<512, 128, true><<<g3, dgmk::wg::NT_SEP, dgmk::wg::SMEM_BYTES_SEP>>>(C, A, nullptr, P, nullptr, 384, 128, M, rps2, 1); }, 10);
        printf("wgrad ZGR warp-specialised, %d splits, no E, separate issuer: %.3f ms  %.2f TFLOP/s\n", splits2, ms, 2.0 * M * 384 * 128 / ms * 1e-9);
        { long long h[32]; CK(cudaMemcpyFromSymbol(h, dgmk::wg::g_wg_prof, sizeof(h))); double n = (double)h[30];
          printf("   cycles/chunk (%d chunks): loader[issue_loads %.0f wait_aempty %.0f split+st %.0f E %.0f] stager[loads|wait_raw %.0f wait_bempty %.0f split+sts %.0f] mma[wait_tempty %.0f wait_afull %.0f wait_bfull %.0f issue %.0f] drain[wait_tfull %.0f drain %.0f]\n",
                 (int)n, h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[8] / n, h[9] / n, h[10] / n, h[18] / n, h[19] / n, h[20] / n, h[21] / n, h[24] / n, h[25] / n); }
        dim3 g4(1, 1, splits2 * 3 > 148 ? 148 : splits2 * 3);
        int64_t rps3 = ((M + g4.z - 1) / g4.z + 31) / 32 * 32;
        ms = time_ms([&] { dgmk::wg::wgrad_ws_kernel<512, 128, true><<<g4, dgmk::wg::NT_SEP, dgmk::wg::SMEM_BYTES_SEP>>>(C, A, nullptr, P, nullptr, 128, 128, M, rps3, 1); }, 10);
        printf("wgrad H   warp-specialised, %d splits, no E, separate issuer: %.3f ms  %.2f TFLOP/s\n", g4.z, ms, 2.0 * M * 128 * 128 / ms * 1e-9);
      }
    }
    CK(cudaGetLastError());
  }
  // ---- tn (weight gradient) tile ----
  {
    CK(cudaFuncSetAttribute(dgmk::tctn::gemm_tn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dgmk::tctn::TN_SMEM_BYTES));
    int64_t M = 1000; int N = 256, Kd = 128; int64_t lda = 512, lds = 128;
    std::vector<float> hA(M * lda), hS(M * lds), hE(M * 4);
    srand(3);
    for (auto& v : hA) v = (rand() / (float)RAND_MAX - 0.5f);
    for (auto& v : hS) v = (rand() / (float)RAND_MAX - 0.5f);
    for (auto& v : hE) v = (rand() / (float)RAND_MAX - 0.5f);
    float *A, *S, *E, *P, *PE;
    int64_t rps = 256; int splits = (int)((M + rps - 1) / rps);
    CK(cudaMalloc(&A, hA.size() * 4)); CK(cudaMalloc(&S, hS.size() * 4)); CK(cudaMalloc(&E, hE.size() * 4));
    CK(cudaMalloc(&P, (size_t)splits * N * Kd * 4)); CK(cudaMalloc(&PE, (size_t)splits * 4 * N * 4));
    CK(cudaMemcpy(A, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(S, hS.data(), hS.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(E, hE.data(), hE.size() * 4, cudaMemcpyHostToDevice));
    dim3 grid(Kd / 128, N / 128, splits);
    {
    dgmk::tctn::gemm_tn_tc_kernel<<<grid, dgmk::tctn::NT, dgmk::tctn::TN_SMEM_BYTES>>>(A, lda, S, lds, E, P, PE, N, Kd, M, rps);
    CK(cudaDeviceSynchronize());
    std::vector<float> hP((size_t)splits * N * Kd), hPE((size_t)splits * 4 * N);
    CK(cudaMemcpy(hP.data(), P, hP.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hPE.data(), PE, hPE.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<float> g((size_t)N * Kd), gr((size_t)N * Kd), ge(4 * N), ger(4 * N);
    for (int n = 0; n < N; ++n) for (int k = 0; k < Kd; ++k) {
      double s = 0, r = 0;
      for (int z = 0; z < splits; ++z) s += hP[((size_t)z * N + n) * Kd + k];
      for (int64_t m = 0; m < M; ++m) r += (double)hA[m * lda + n] * hS[m * lds + k];
      g[(size_t)n * Kd + k] = (float)s; gr[(size_t)n * Kd + k] = (float)r;
    }
    for (int e = 0; e < 4; ++e) for (int n = 0; n < N; ++n) {
      double s = 0, r = 0;
      for (int z = 0; z < splits; ++z) s += hPE[((size_t)z * 4 + e) * N + n];
      for (int64_t m = 0; m < M; ++m) r += (double)hA[m * lda + n] * hE[m * 4 + e];
      ge[e * N + n] = (float)s; ger[e * N + n] = (float)r;
    }
    printf("gemm_tn_tc M=%ld N=%d Kd=%d relerr W %.3e  E %.3e   g[0..3]= %g %g %g %g ref %g %g %g %g\n", (long)M, N, Kd, relerr(g, gr), relerr(ge, ger),
           g[0], g[1], g[2], g[3], gr[0], gr[1], gr[2], gr[3]);
    printf("   g[128*Kd..]= %g %g ref %g %g ; ge[0..1]= %g %g ref %g %g\n", g[128 * Kd], g[128 * Kd + 1], gr[128 * Kd], gr[128 * Kd + 1], ge[0], ge[1], ger[0], ger[1]);
    {
      CK(cudaFuncSetAttribute(dgmk::wg::wgrad_ws_kernel<512, 128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dgmk::wg::SMEM_BYTES));
      CK(cudaFuncSetAttribute(dgmk::wg::wgrad_ws_kernel<512, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dgmk::wg::SMEM_BYTES_SEP));
      // two segments of 128 rows per CTA -> 2*splits W partials, 4*splits E partials
      cudaFree(P); CK(cudaMalloc(&P, (size_t)2 * splits * N * Kd * 4));
      float* PE2; CK(cudaMalloc(&PE2, (size_t)4 * splits * 4 * N * 4));
      CK(cudaMemset(P, 0, (size_t)2 * splits * N * Kd * 4)); CK(cudaMemset(PE2, 0, (size_t)4 * splits * 4 * N * 4));
      hP.assign((size_t)2 * splits * N * Kd, 0.f); hPE.assign((size_t)4 * splits * 4 * N, 0.f);
      dgmk::wg::wgrad_ws_kernel<512, 128, false><<<grid, dgmk::wg::NT, dgmk::wg::SMEM_BYTES>>>(A, S, E, P, PE2, N, Kd, M, rps / 2, 2);
      CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(hP.data(), P, hP.size() * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hPE.data(), PE2, hPE.size() * 4, cudaMemcpyDeviceToHost));
      for (int n = 0; n < N; ++n) for (int k = 0; k < Kd; ++k) { double s = 0; for (int z = 0; z < 2 * splits; ++z) s += hP[((size_t)z * N + n) * Kd + k]; g[(size_t)n * Kd + k] = (float)s; }
      for (int e = 0; e < 4; ++e) for (int n = 0; n < N; ++n) { double s = 0; for (int z = 0; z < 4 * splits; ++z) s += hPE[((size_t)z * 4 + e) * N + n]; ge[e * N + n] = (float)s; }
      printf("wgrad_ws   M=%ld N=%d Kd=%d relerr W %.3e  E %.3e\n", (long)M, N, Kd, relerr(g, gr), relerr(ge, ger));
      // the variant with its own MMA-issuer warpgroup (no E)
      CK(cudaMemset(P, 0, (size_t)2 * splits * N * Kd * 4));
      dgmk::wg::wgrad_ws_kernel<512, 128, true><<<grid, dgmk::wg::NT_SEP, dgmk::wg::SMEM_BYTES_SEP>>>(A, S, nullptr, P, nullptr, N, Kd, M, rps / 2, 2);
      CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(hP.data(), P, hP.size() * 4, cudaMemcpyDeviceToHost));
      for (int n = 0; n < N; ++n) for (int k = 0; k < Kd; ++k) { double s = 0; for (int z = 0; z < 2 * splits; ++z) s += hP[((size_t)z * N + n) * Kd + k]; g[(size_t)n * Kd + k] = (float)s; }
      printf("wgrad_ws (separate issuer) M=%ld N=%d Kd=%d relerr W %.3e\n", (long)M, N, Kd, relerr(g, gr));
    }
    }
  }
  printf("done\n");
  return 0;
}
