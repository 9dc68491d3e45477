// Microbenchmarks run on the B200 box (gpurun): FP32 FFMA peak (the roofline
// denominator MEASURED_PEAKS.json lacks) and the stand-alone GEMM
// tiles of csrc/dgmk_gemm.cuh (correctness vs a naive kernel + TFLOP/s).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o bench_gemm bench_gemm.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../differential_equations_dnn_b200/csrc/dgmk_gemm.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// outer-product FFMA: 8x8 accumulators, operands in registers (same shape as the
// SGEMM inner loop, so ptxas can use the operand-reuse cache).
__global__ void __launch_bounds__(256) ffma_outer(float* out, const float* in, int iters) {
  float a[8], b[8], acc[8][8];
  for (int i = 0; i < 8; ++i) { a[i] = in[threadIdx.x % 32 + i]; b[i] = in[8 + threadIdx.x % 16 + i]; }
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// dependent-chain form: acc = acc * b + c, 16 independent chains
__global__ void __launch_bounds__(256) ffma_chain(float* out, const float* in, int iters) {
  float acc[16];
  float b = in[threadIdx.x % 7], c = in[threadIdx.x % 5 + 1];
  for (int i = 0; i < 16; ++i) acc[i] = in[i];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], b, c);
  }
  float s = 0.f;
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void naive_nn(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                         int64_t M, int N, int K, bool accum) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * N) return;
  int64_t m = idx / N; int n = idx % N;
  double s = accum ? C[m * ldc + n] : 0.0;
  for (int k = 0; k < K; ++k) s += (double)A[m * lda + k] * B[(int64_t)k * ldb + n];
  C[m * ldc + n] = (float)s;
}
__global__ void naive_tn(const float* A, int64_t lda, const float* S, int64_t lds, float* G, int N, int Kd, int64_t M) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * Kd) return;
  int n = idx / Kd, k = idx % Kd;
  double s = 0;
  for (int64_t m = 0; m < M; ++m) s += (double)A[m * lda + n] * S[m * lds + k];
  G[idx] = (float)s;
}
__global__ void sum_parts(const float* P, int splits, int n, float* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0;
  for (int z = 0; z < splits; ++z) s += P[(int64_t)z * n + i];
  out[i] = (float)s;
}

static double relerr(const std::vector<float>& a, const std::vector<float>& b) {
  double num = 0, den = 0;
  for (size_t i = 0; i < a.size(); ++i) { double d = (double)a[i] - b[i]; num += d * d; den += (double)b[i] * b[i]; }
  return sqrt(num / (den + 1e-300));
}

template <typename F>
float time_ms(F f, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("device %s SMs %d clock %d kHz\n", prop.name, sms, prop.clockRate);
  float *in, *out;
  CK(cudaMalloc(&in, 4096)); CK(cudaMalloc(&out, sizeof(float) * sms * 16 * 256));
  std::vector<float> hin(1024, 1.0f); for (int i = 0; i < 1024; ++i) hin[i] = 1.0f + 1e-7f * i;
  CK(cudaMemcpy(in, hin.data(), 4096, cudaMemcpyHostToDevice));
  for (int bps : {2, 4, 8}) {
    int iters = 20000;
    int grid = sms * bps;
    float ms = time_ms([&] { ffma_outer<<<grid, 256>>>(out, in, iters); }, 5);
    double fl = 2.0 * 64 * iters * 256.0 * grid;
    printf("ffma_outer blocks/SM %d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
    ms = time_ms([&] { ffma_chain<<<grid, 256>>>(out, in, iters); }, 5);
    fl = 2.0 * 64 * iters * 256.0 * grid;
    printf("ffma_chain blocks/SM %d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
  }
  // sustained (about 2 s) outer-product figure for a kernel inside a long step
  {
    int iters = 20000, grid = sms * 4;
    float one = time_ms([&] { ffma_outer<<<grid, 256>>>(out, in, iters); }, 3);
    int reps = (int)(2000.0f / one) + 1;
    float ms = time_ms([&] { ffma_outer<<<grid, 256>>>(out, in, iters); }, reps);
    double fl = 2.0 * 64 * iters * 256.0 * grid;
    printf("ffma_outer sustained (%d reps): %.3f ms  %.2f TFLOP/s\n", reps, ms, fl / ms * 1e-9);
  }

  // ---- GEMM tiles ----
  struct Case { int64_t M; int N, K; bool accum; };
  Case cases[] = {{1000, 384, 128, false}, {777, 128, 384, true}, {513, 64, 64, false}, {300, 32, 32, true},
                  {300, 96, 32, false}};
  for (auto c : cases) {
    int64_t lda = c.K + 32, ldc = c.N + 64;
    std::vector<float> hA(c.M * lda), hB((size_t)c.K * c.N), hC(c.M * ldc);
    srand(1);
    for (auto& v : hA) v = (rand() / (float)RAND_MAX - 0.5f);
    for (auto& v : hB) v = (rand() / (float)RAND_MAX - 0.5f);
    for (auto& v : hC) v = (rand() / (float)RAND_MAX - 0.5f);
    float *A, *B, *C, *Cr;
    CK(cudaMalloc(&A, hA.size() * 4)); CK(cudaMalloc(&B, hB.size() * 4));
    CK(cudaMalloc(&C, hC.size() * 4)); CK(cudaMalloc(&Cr, hC.size() * 4));
    CK(cudaMemcpy(A, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(B, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(C, hC.data(), hC.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(Cr, hC.data(), hC.size() * 4, cudaMemcpyHostToDevice));
    naive_nn<<<(unsigned)((c.M * c.N + 255) / 256), 256>>>(A, lda, B, c.N, Cr, ldc, c.M, c.N, c.K, c.accum);
    int BN = (c.N % 128 == 0) ? 128 : (c.N % 64 == 0) ? 64 : 32;
    dim3 grid(c.N / BN, (unsigned)((c.M + 127) / 128));
#define LAUNCH_NN(bn) \
    if (c.accum) dgmk::gemm_nn_kernel<bn, true><<<grid, 256>>>(A, lda, B, c.N, C, ldc, c.M, c.K); \
    else dgmk::gemm_nn_kernel<bn, false><<<grid, 256>>>(A, lda, B, c.N, C, ldc, c.M, c.K);
    if (BN == 128) { LAUNCH_NN(128) } else if (BN == 64) { LAUNCH_NN(64) } else { LAUNCH_NN(32) }
    CK(cudaDeviceSynchronize());
    std::vector<float> r1(hC.size()), r2(hC.size());
    CK(cudaMemcpy(r1.data(), C, hC.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(r2.data(), Cr, hC.size() * 4, cudaMemcpyDeviceToHost));
    printf("gemm_nn M=%ld N=%d K=%d accum=%d BN=%d relerr %.3e\n", (long)c.M, c.N, c.K, c.accum, BN, relerr(r1, r2));
    // tn: G[N,K] from A'[M,N] (reuse C as A') and S[M,K] (reuse A)
    {
      int Kd = c.K, N = c.N; int64_t rps = 256; int splits = (int)((c.M + rps - 1) / rps);
      float *P, *G, *Gr;
      CK(cudaMalloc(&P, (size_t)splits * N * Kd * 4)); CK(cudaMalloc(&G, N * Kd * 4)); CK(cudaMalloc(&Gr, N * Kd * 4));
      int BNt = (Kd % 128 == 0) ? 128 : (Kd % 64 == 0) ? 64 : 32;
      dim3 g2(Kd / BNt, (N + 127) / 128, splits);
      if (BNt == 128) dgmk::gemm_tn_kernel<128><<<g2, 256>>>(Cr, ldc, A, lda, P, N, Kd, c.M, rps, nullptr, nullptr);
      else if (BNt == 64) dgmk::gemm_tn_kernel<64><<<g2, 256>>>(Cr, ldc, A, lda, P, N, Kd, c.M, rps, nullptr, nullptr);
      else dgmk::gemm_tn_kernel<32><<<g2, 256>>>(Cr, ldc, A, lda, P, N, Kd, c.M, rps, nullptr, nullptr);
      sum_parts<<<(N * Kd + 255) / 256, 256>>>(P, splits, N * Kd, G);
      naive_tn<<<(N * Kd + 255) / 256, 256>>>(Cr, ldc, A, lda, Gr, N, Kd, c.M);
      CK(cudaDeviceSynchronize());
      std::vector<float> g1(N * Kd), g2v(N * Kd);
      CK(cudaMemcpy(g1.data(), G, g1.size() * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(g2v.data(), Gr, g1.size() * 4, cudaMemcpyDeviceToHost));
      printf("gemm_tn M=%ld N=%d Kd=%d BN=%d relerr %.3e\n", (long)c.M, N, Kd, BNt, relerr(g1, g2v));
      cudaFree(P); cudaFree(G); cudaFree(Gr);
    }
    cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(Cr);
  }
  // ---- throughput at the heat/DGM(128,3) shapes: M = 4 * 2^17 jet rows ----
  {
    int64_t M = 4LL << 17;
    float *A, *B, *C, *P;
    CK(cudaMalloc(&A, M * 512 * 4)); CK(cudaMalloc(&B, 512 * 512 * 4)); CK(cudaMalloc(&C, M * 512 * 4));
    CK(cudaMemset(A, 0, M * 512 * 4)); CK(cudaMemset(B, 0, 512 * 512 * 4));
    int64_t rps = 4096; int splits = (int)(M / rps);
    CK(cudaMalloc(&P, (size_t)splits * 384 * 128 * 4));
    struct T { const char* name; int N, K; bool acc; } ts[] = {
        {"fwd ZGR  [M,128]x[128,384]", 384, 128, false}, {"fwd H    [M,128]x[128,128]", 128, 128, false},
        {"dgrad ZGR[M,384]x[384,128]", 128, 384, true}};
    for (auto t : ts) {
      dim3 grid(t.N / 128, (unsigned)(M / 128));
      float ms = time_ms([&] {
        if (t.acc) dgmk::gemm_nn_kernel<128, true><<<grid, 256>>>(A, 512, B, t.N, C, 512, M, t.K);
        else dgmk::gemm_nn_kernel<128, false><<<grid, 256>>>(A, 512, B, t.N, C, 512, M, t.K);
      }, 10);
      printf("%s: %.3f ms  %.2f TFLOP/s\n", t.name, ms, 2.0 * M * t.N * t.K / ms * 1e-9);
    }
    dim3 g2(1, 3, splits);
    float ms = time_ms([&] { dgmk::gemm_tn_kernel<128><<<g2, 256>>>(C, 512, A, 512, P, 384, 128, M, rps, nullptr, nullptr); }, 10);
    printf("wgrad ZGR [M,384]^T x [M,128]: %.3f ms  %.2f TFLOP/s\n", ms, 2.0 * M * 384 * 128 / ms * 1e-9);
    CK(cudaGetLastError());
  }
  printf("done\n");
  return 0;
}
