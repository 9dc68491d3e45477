// FFMA2 (fma.rn.f32x2, sm_100+) probes: is the packed form able to beat the 3-register
// FFMA outer product (60.5 TFLOP/s measured) that bounds the SGEMM inner loop?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// outer product, packed along columns: acc2[i][jj] += (a_i, a_i) * (b_2jj, b_2jj+1)
__global__ void __launch_bounds__(256) ffma2_outer_dupA(float* out, const float* in, int iters) {
  float a[8]; float2 b2[4], acc[8][4];
  for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x % 32 + i];
  for (int j = 0; j < 4; ++j) b2[j] = make_float2(in[8 + threadIdx.x % 16 + 2 * j], in[9 + threadIdx.x % 16 + 2 * j]);
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float2 a2 = make_float2(a[i], a[i]);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = __ffma2_rn(a2, b2[j], acc[i][j]);
    }
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j].x + acc[i][j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// same, but the duplicated pairs are prepared outside the loop (upper bound: no MOVs)
__global__ void __launch_bounds__(256) ffma2_outer_predup(float* out, const float* in, int iters) {
  float2 a2[8], b2[4], acc[8][4];
  for (int i = 0; i < 8; ++i) { float v = in[threadIdx.x % 32 + i]; a2[i] = make_float2(v, v); }
  for (int j = 0; j < 4; ++j) b2[j] = make_float2(in[8 + threadIdx.x % 16 + 2 * j], in[9 + threadIdx.x % 16 + 2 * j]);
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = __ffma2_rn(a2[i], b2[j], acc[i][j]);
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j].x + acc[i][j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// chain form with packed operands
__global__ void __launch_bounds__(256) ffma2_chain(float* out, const float* in, int iters) {
  float2 acc[8];
  float2 b = make_float2(in[threadIdx.x % 7], in[threadIdx.x % 7 + 1]), c = make_float2(in[threadIdx.x % 5 + 1], in[threadIdx.x % 5]);
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(in[i], in[i + 1]);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = __ffma2_rn(acc[i], b, c);
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 3-register scalar FFMA outer product (reference point, 60.5 TF in the first run)
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

template <typename F> float time_ms(F f, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) f();
  CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms / reps;
}
int main() {
  int sms = 148; float *in, *out;
  CK(cudaMalloc(&in, 4096)); CK(cudaMalloc(&out, sizeof(float) * sms * 16 * 256));
  float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = 1.0f + 1e-7f * i;
  CK(cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice));
  int iters = 20000;
  for (int bps : {1, 2, 4}) {
    int grid = sms * bps; double fl = 2.0 * 64 * iters * 256.0 * grid;
    float ms = time_ms([&] { ffma_outer<<<grid, 256>>>(out, in, iters); }, 5);
    printf("blocks/SM %d  ffma_outer        %.2f TFLOP/s\n", bps, fl / ms * 1e-9);
    ms = time_ms([&] { ffma2_outer_dupA<<<grid, 256>>>(out, in, iters); }, 5);
    printf("blocks/SM %d  ffma2_outer_dupA  %.2f TFLOP/s\n", bps, fl / ms * 1e-9);
    ms = time_ms([&] { ffma2_outer_predup<<<grid, 256>>>(out, in, iters); }, 5);
    printf("blocks/SM %d  ffma2_outer_predup %.2f TFLOP/s\n", bps, fl / ms * 1e-9);
    ms = time_ms([&] { ffma2_chain<<<grid, 256>>>(out, in, iters); }, 5);
    printf("blocks/SM %d  ffma2_chain       %.2f TFLOP/s\n", bps, fl / ms * 1e-9);
  }
  return 0;
}
