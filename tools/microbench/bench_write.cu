// HBM write / read / copy bandwidth with different store shapes (calibration for the fused epilogues)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
template <typename F> float time_ms(F f, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 2; ++i) f();
  CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms / reps;
}
// each thread stores VEC floats; warp covers 32*VEC contiguous floats; grid-stride
template <int VEC>
__global__ void fill(float* p, int64_t n) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC, stride = (int64_t)gridDim.x * blockDim.x * VEC;
  for (; i < n; i += stride) {
    if (VEC == 4) *reinterpret_cast<float4*>(p + i) = make_float4(1.f, 2.f, 3.f, 4.f);
    else p[i] = 1.f;
  }
}
__global__ void readk(const float4* p, int64_t n, float* out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  float s = 0.f;
  for (; i < n; i += stride) { float4 v = p[i]; s += v.x + v.y + v.z + v.w; }
  if (s == 1234.5f) *out = s;
}
// persistent: CTA b owns tiles b, b+grid, ...; tile = 64 rows x 128 floats of a [M, ld] matrix at column block g;
// 8 warps: warp w stores rows (32 per half) one 128-byte piece per instruction: the lane-GEMM epilogue pattern
__global__ void __launch_bounds__(256) tile_store(float* C, int64_t ld, int64_t M, int ngates) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gate = blockIdx.x % ngates; const int64_t grp = blockIdx.x / ngates, ngrp = gridDim.x / ngates;
  const int quarter = warp & 3, half = warp >> 2;
  for (int64_t t = grp; t * 64 < M; t += ngrp) {
    float* c = C + (t * 64 + half * 32) * ld + gate * 128 + quarter * 32 + lane;
#pragma unroll 8
    for (int q = 0; q < 32; ++q) c[q * ld] = (float)q;
  }
}
int main() {
  int64_t n = 1LL << 28;  // 1 GiB of floats
  float* p; CK(cudaMalloc(&p, n * 4)); float* out; CK(cudaMalloc(&out, 4));
  float ms;
  ms = time_ms([&] { CK(cudaMemsetAsync(p, 0, n * 4)); }, 5); printf("cudaMemset 1 GiB: %.3f ms  %.0f GB/s\n", ms, n * 4 / ms * 1e-6);
  ms = time_ms([&] { fill<4><<<148 * 16, 256>>>(p, n); }, 5); printf("fill float4 grid-stride: %.3f ms  %.0f GB/s\n", ms, n * 4 / ms * 1e-6);
  ms = time_ms([&] { fill<1><<<148 * 16, 256>>>(p, n); }, 5); printf("fill float  grid-stride: %.3f ms  %.0f GB/s\n", ms, n * 4 / ms * 1e-6);
  ms = time_ms([&] { readk<<<148 * 16, 256>>>((const float4*)p, n / 4, out); }, 5); printf("read float4 grid-stride: %.3f ms  %.0f GB/s\n", ms, n * 4 / ms * 1e-6);
  int64_t M = 1LL << 19;
  for (int ng : {3, 1}) for (int64_t ld : {512, 384, 128}) {
    if (ld < ng * 128) continue;
    for (int ctas : {1, 2, 4}) {
      int grid = 148 * ctas / ng * ng;
      ms = time_ms([&] { tile_store<<<grid, 256>>>(p, ld, M, ng); }, 5);
      printf("tile_store ngates=%d ld=%ld ctas/SM=%d: %.3f ms  %.0f GB/s\n", ng, (long)ld, ctas, ms, (double)M * ng * 512 / ms * 1e-6);
    }
  }
  return 0;
}
