// CUDA backend (sm_100a) + the extern "C" surface of include/dgmk.h.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
//        --expt-relaxed-constexpr -Xcompiler -fPIC -shared -o libdgmk.so dgmk_cuda.cu
// This is the only implementation the package loads: there is no CPU path.
#include <cuda_runtime.h>
#include <atomic>
#include <cstdlib>
#include <vector>
#include "dgmk_capi_impl.h"
#include "dgmk_gemm.cuh"
#include "dgmk_gemm_tc.cuh"
#include "dgmk_gemm_tc_tn.cuh"
#include "dgmk_wgrad_ws.cuh"
#include "dgmk_lane_gemm.cuh"
#include "dgmk_lane_epi.cuh"
#include "dgmk_dgrad_res.cuh"
#include "dgmk_tile_params.h"

namespace dgmk {

constexpr int EW_THREADS = 256;

// number of kernels this library has launched in this process (diagnostic: bench.py
// reports it as gpu_launches)
static unsigned long long g_launches = 0;
// GEMM engine selector for A/B measurements (dgmk_set_gemm_engine): tensor cores on by default
static bool g_use_tc = true;
// fused units-on-lanes kernels (GEMM + element-wise stage in one launch) where the shape allows
static bool g_fuse = true;
// K = 3H data gradient on the weight-resident kernel (dgmk_dgrad_res.cuh); off: the round-1 streaming tile
static bool g_dgrad_res = true;
// resident-tile step (one persistent kernel per step) for hidden sizes <= 64 (dgmk_set_tile_engine)
static int g_tile = 1;   // 0 off, 1 on where it wins (dispatch rule in tile_step), 2 forced on wherever it fits
static long long* g_tile_prof = nullptr;   // dgmk_tile_profile: device buffer for CTA 0's stage timeline
static int g_tile_prof_n = 0;
static int g_tile_flush = tk::FLUSH_TILES;   // dgmk_set_tile_flush (tests: several accumulation segments on small batches)
namespace tk { int launch(int prob, const TileParams& prm, int grid, size_t smem, void* stream); }   // dgmk_tile.cu

// ---- per-kernel-class timing (dgmk_profile*): CUDA events recorded on the launch stream around
// every launch of a class, with the launch's ALGORITHMIC flops and bytes, so that bench.py can
// report the roofline of the dominant kernel from the timed region itself ----------------------
enum { PC_WGRAD = 0, PC_LANE = 1, PC_STREAM_NN = 2, PC_EW = 3, PC_OTHER = 4, PC_TILE = 5, PC_COUNT = 6 };
struct ProfClass {
  std::vector<cudaEvent_t> ev;   // begin, end, begin, end, ...
  double flops = 0.0, bytes = 0.0;
  long long launches = 0;
};
static bool g_prof_on = false;
static ProfClass g_prof[PC_COUNT];
static std::vector<cudaEvent_t> g_prof_pool;
static cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
struct ProfScope {   // records begin now, end at destruction
  int cls; cudaStream_t st; bool on;
  ProfScope(int c, cudaStream_t s, double flops, double bytes) : cls(c), st(s), on(g_prof_on) {
    if (!on) return;
    cudaEvent_t e = prof_event();
    cudaEventRecord(e, st);
    g_prof[cls].ev.push_back(e);
    g_prof[cls].flops += flops; g_prof[cls].bytes += bytes; g_prof[cls].launches += 1;
  }
  ~ProfScope() {
    if (!on) return;
    cudaEvent_t e = prof_event();
    cudaEventRecord(e, st);
    g_prof[cls].ev.push_back(e);
  }
};

template <class F>
__global__ void __launch_bounds__(EW_THREADS) ew_kernel(const F f, int64_t n) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}

// functors with a vec4(k) method: one thread = four consecutive hidden units of one row
template <class F>
__global__ void __launch_bounds__(EW_THREADS) ew4_kernel(const F f, int64_t n4) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n4; k += stride) f.vec4(k);
}

// deterministic second stage: out[(i / cols) * ldo + i % cols] += sum_p part[p][i], i < n.  FP64, fixed
// order: blockDim = (32 outputs, S slices of the partials); a thread adds its slice with four
// independent chains (loads in flight), the slices are combined in slice order.  (One thread per
// output walking all partials -- the first version -- was latency-bound: 50-260 us per launch.)
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const float* __restrict__ part, int nparts, int64_t n,
                                                               float* __restrict__ out, int cols, int64_t ldo, bool overwrite) {
  const int64_t i = (int64_t)blockIdx.x * 32 + threadIdx.x;
  const int S = blockDim.y, sl = threadIdx.y;
  __shared__ double sm[32][33];
  double s = 0.0;
  if (i < n) {
    const int chunk = (nparts + S - 1) / S;
    const int p0 = sl * chunk, p1 = (p0 + chunk < nparts) ? p0 + chunk : nparts;
    const float* q = part + (int64_t)p0 * n + i;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int p = p0;
    for (; p + 4 <= p1; p += 4, q += 4 * n) {
      const float v0 = __ldg(q), v1 = __ldg(q + n), v2 = __ldg(q + 2 * n), v3 = __ldg(q + 3 * n);
      s0 += (double)v0; s1 += (double)v1; s2 += (double)v2; s3 += (double)v3;
    }
    for (; p < p1; ++p, q += n) s0 += (double)__ldg(q);
    s = (s0 + s1) + (s2 + s3);
  }
  sm[sl][threadIdx.x] = s;
  __syncthreads();
  if (sl == 0 && i < n) {
    double t = 0.0;
    for (int k = 0; k < S; ++k) t += sm[k][threadIdx.x];
    const int64_t r = i / cols, c = i - r * cols;
    out[r * ldo + c] = overwrite ? (float)t : out[r * ldo + c] + (float)t;
  }
}

// Resident-tile step, small batches: second-stage reduction of the per-CTA partial gradients, scatter into the
// caller's flat gradient (UnpackGradFn's mapping) and the loss in ONE launch.  Thread i < P owns parameter i and adds
// its packed element over the nparts slots in FP64 (four chains, slot order); thread P does the loss.
__global__ void __launch_bounds__(128) tile_unpack_kernel(const SegTable t, const float* __restrict__ part, int nparts, int64_t g_floats,
                                                          int64_t g_acc, float* __restrict__ grad, float* __restrict__ loss, int64_t P) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > P) return;
  int64_t off = -1;
  if (i == P) off = g_acc;
  else {
    for (int s = 0; s < t.n; ++s) {
      const Seg& g = t.s[s];
      const int32_t k = (int32_t)i - g.theta_off;
      if (k >= 0 && k < g.n) {
        const int r = k / g.cols, c = k - r * g.cols;
        if (g.a_off >= 0) off = g.a_off + (int64_t)r * g.a_rs + (int64_t)c * g.a_cs;
        break;
      }
    }
  }
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (off >= 0) {
    const float* q = part + off;
    int p = 0;
    for (; p + 4 <= nparts; p += 4, q += 4 * g_floats) {
      const float v0 = __ldg(q), v1 = __ldg(q + g_floats), v2 = __ldg(q + 2 * g_floats), v3 = __ldg(q + 3 * g_floats);
      s0 += (double)v0; s1 += (double)v1; s2 += (double)v2; s3 += (double)v3;
    }
    for (; p < nparts; ++p, q += g_floats) s0 += (double)__ldg(q);
  }
  const float v = (float)((s0 + s1) + (s2 + s3));
  if (i == P) { if (loss) *loss = v; }
  else if (grad) grad[i] = v;
}

// DgmRev1Fn (hidden size 128) that also forms grad[U | b] of the Z, G and H gates -- the input-map
// adjoint of the three pre-activation cotangents it has just computed -- so that the weight-gradient
// kernel has no A^T E work left.  part[blk][gate 0..2 = Z, G, H][e 0..2 = U[:,0], U[:,1], b][128]
// V units per thread (8- / 16-byte accesses): with one unit per thread the value-only stage ran at 0.74 of
// the measured HBM peak; V = 4 / 2 took 3.8 ms off the step.  The grid stride is a multiple of 128 / V: a
// thread keeps its unit group q = tid % (128 / V) and nine running sums per unit for its whole life; 2 V
// threads of a block share a group.
template <class CS, int V>
struct Rev1SinkV {
  float x0, x1; float (*g)[3][3];
  __device__ __forceinline__ void operator()(int u, int slot, const float* ab) const { input_map_adj<CS>(ab, x0, x1, g[u][slot == 3 ? 2 : slot]); }
};
template <class F, class CS, int V, int MINB>
__global__ void __launch_bounds__(EW_THREADS, MINB) rev1_ev_kernel(const F f, const XSrc xs, int64_t nv, float* __restrict__ part) {
  constexpr int GROUPS = 128 / V;
  float g[V][3][3];
#pragma unroll
  for (int u = 0; u < V; ++u)
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int e = 0; e < 3; ++e) g[u][a][e] = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nv; k += stride) {
    const float* x = xs.at(k / GROUPS);
    Rev1SinkV<CS, V> sink;
    sink.x0 = __ldg(x); sink.x1 = (xs.d > 1) ? __ldg(x + 1) : 0.f; sink.g = g;
    f.template runv<V>(k, sink);
  }
  __shared__ float sm[V * 9][EW_THREADS];
#pragma unroll
  for (int u = 0; u < V; ++u)
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int e = 0; e < 3; ++e) sm[u * 9 + a * 3 + e][threadIdx.x] = g[u][a][e];
  __syncthreads();
  if (threadIdx.x < 128) {
    const int q = threadIdx.x / V, u = threadIdx.x % V;
#pragma unroll
    for (int ge = 0; ge < 9; ++ge) {
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < EW_THREADS / GROUPS; ++r) s += sm[u * 9 + ge][q + GROUPS * r];
      part[((int64_t)blockIdx.x * 9 + ge) * 128 + threadIdx.x] = s;
    }
  }
}

// second stage of the fused input-map gradients: part[p][g][e][128] -> out[e * ldo + gmap[g] * 128 + j] +=
// sum_p, FP64, fixed order (slices of the partials per output, combined in slice order)
__global__ void __launch_bounds__(1024) reduce_gate_e_kernel(const float* __restrict__ part, int nparts, int ngates, int gm0,
                                                             int gm1, int gm2, float* __restrict__ out, int64_t ldo) {
  const int o = blockIdx.x * 32 + threadIdx.x, S = blockDim.y, sl = threadIdx.y;
  const int per = ngates * 384;
  __shared__ double sm[32][33];
  double s = 0.0;
  if (o < per) {
    const int chunk = (nparts + S - 1) / S;
    const int p0 = sl * chunk, p1 = (p0 + chunk < nparts) ? p0 + chunk : nparts;
    const float* q = part + (int64_t)p0 * per + o;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int p = p0;
    for (; p + 4 <= p1; p += 4, q += 4 * per) {
      const float v0 = __ldg(q), v1 = __ldg(q + per), v2 = __ldg(q + 2 * per), v3 = __ldg(q + 3 * per);
      s0 += (double)v0; s1 += (double)v1; s2 += (double)v2; s3 += (double)v3;
    }
    for (; p < p1; ++p, q += per) s0 += (double)__ldg(q);
    s = (s0 + s1) + (s2 + s3);
  }
  sm[sl][threadIdx.x] = s;
  __syncthreads();
  if (sl == 0 && o < per) {
    double t = 0.0;
    for (int k = 0; k < S; ++k) t += sm[k][threadIdx.x];
    const int g = o / 384, e = (o / 128) % 3, j = o & 127;
    const int gm = (g == 0) ? gm0 : (g == 1 ? gm1 : gm2);
    out[(int64_t)e * ldo + gm * 128 + j] += (float)t;
  }
}

// out-partials[blk][e][n] = sum_{r in block's row strip} Wt[r][e] * Mat[r][n]
// grid = (ceil(N / 128), nblk); thread = one column n, 4 weighted sums.
template <bool WEIGHTED>
__global__ void __launch_bounds__(128) wcolsum_kernel(const float* __restrict__ Mat, int64_t ldm, int N,
                                                      const float* __restrict__ Wt, int64_t M, int64_t rows_per_blk,
                                                      float* __restrict__ part) {
  const int n = blockIdx.x * 128 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_blk;
  const int64_t r1 = (r0 + rows_per_blk < M) ? r0 + rows_per_blk : M;
  constexpr int NE = WEIGHTED ? 4 : 1;
  float acc[NE];
#pragma unroll
  for (int e = 0; e < NE; ++e) acc[e] = 0.f;
  __shared__ float4 wsm[128];
  for (int64_t rb = r0; rb < r1; rb += 128) {
    int cnt = (int)((r1 - rb < 128) ? r1 - rb : 128);
    if (WEIGHTED) {
      __syncthreads();
      if (threadIdx.x < cnt) wsm[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(Wt) + rb + threadIdx.x);
      __syncthreads();
    }
    if (n < N) {
      int q = 0;
      for (; q + 8 <= cnt; q += 8) {   // 8 independent loads in flight per thread
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(Mat + (rb + q + u) * ldm + n);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (WEIGHTED) {
            float4 w = wsm[q + u];
            acc[0] = fmaf(w.x, v[u], acc[0]);
            acc[1] = fmaf(w.y, v[u], acc[1]);
            acc[2] = fmaf(w.z, v[u], acc[2]);
            acc[NE - 1] = fmaf(w.w, v[u], acc[NE - 1]);
          } else {
            acc[0] += v[u];
          }
        }
      }
      for (; q < cnt; ++q) {
        float v = __ldg(Mat + (rb + q) * ldm + n);
        if (WEIGHTED) {
          float4 w = wsm[q];
          acc[0] = fmaf(w.x, v, acc[0]);
          acc[1] = fmaf(w.y, v, acc[1]);
          acc[2] = fmaf(w.z, v, acc[2]);
          acc[NE - 1] = fmaf(w.w, v, acc[NE - 1]);
        } else {
          acc[0] += v;
        }
      }
    }
  }
  if (n < N) {
#pragma unroll
    for (int e = 0; e < NE; ++e) part[((int64_t)blockIdx.y * NE + e) * N + n] = acc[e];
  }
}

// u[r][m] = S[r,:] . W[m,:] (+ b[m] on value rows); one warp per row
__global__ void __launch_bounds__(256) rowdot_kernel(const float* __restrict__ S, int64_t lds, const float* __restrict__ W,
                                                     const float* __restrict__ b, float* __restrict__ U, int64_t M, int Hp,
                                                     int o, int C) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < M; r += nwarps) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = lane; j < Hp; j += 32) {
      float s = __ldg(S + r * lds + j);
#pragma unroll
      for (int m = 0; m < 4; ++m)
        if (m < o) acc[m] = fmaf(s, __ldg(W + m * Hp + j), acc[m]);
    }
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], off);
    if (lane == 0) {
      const bool vrow = (r % C) == 0;
      float4 out;
      out.x = acc[0] + ((vrow && 0 < o) ? __ldg(b + 0) : 0.f);
      out.y = (1 < o) ? acc[1] + (vrow ? __ldg(b + 1) : 0.f) : 0.f;
      out.z = (2 < o) ? acc[2] + (vrow ? __ldg(b + 2) : 0.f) : 0.f;
      out.w = (3 < o) ? acc[3] + (vrow ? __ldg(b + 3) : 0.f) : 0.f;
      *reinterpret_cast<float4*>(U + r * 4) = out;
    }
  }
}

struct CudaBackend : BackendTraitsAll {
  cudaStream_t st;
  const char* err;
  int sms;
  bool use_tc, fuse, dgrad_res;
  int64_t hl_stride = 0;  // distance between the plain / tf32-hi / tf32-lo copies of the packed weights
  // design bytes (operands read + results written, each once) of the NEXT element-wise / reduction launch,
  // announced by the pipeline for the per-class traffic accounting of dgmk_profile
  double pending_bytes = 0.0;
  void note_bytes(double b) { pending_bytes = b; }
  double take_bytes() { double b = pending_bytes; pending_bytes = 0.0; return b; }
  explicit CudaBackend(void* stream) : st((cudaStream_t)stream), err(nullptr), sms(148), use_tc(g_use_tc), fuse(g_fuse), dgrad_res(g_dgrad_res) {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) {
      int v = 0;
      if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
    }
  }
  void note(cudaError_t e) { if (e != cudaSuccess && !err) err = cudaGetErrorString(e); }
  // true exactly once per (kernel family `slot`, current device): cudaFuncSetAttribute opt-ins are per device.
  // Setting the attribute twice is harmless, so a relaxed race between threads only repeats the call.
  bool first_use_on_device(int slot) {
    static std::atomic<unsigned long long> masks[8];
    int dev = 0;
    note(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (masks[slot].load(std::memory_order_acquire) & bit) return false;
    masks[slot].fetch_or(bit, std::memory_order_release);
    return true;
  }
  void post() { ++g_launches; note(cudaPeekAtLastError()); }

  template <class F>
  void ew(const F& f, int64_t n) {
    if (n <= 0) return;
    int64_t blocks = (n + EW_THREADS - 1) / EW_THREADS;
    int64_t cap = (int64_t)sms * 32;
    if (blocks > cap) blocks = cap;
    ProfScope ps(PC_EW, st, 0.0, take_bytes());
    ew_kernel<F><<<(unsigned)blocks, EW_THREADS, 0, st>>>(f, n);
    post();
  }
  // n = rows * Hp elements, Hp a multiple of 32: four units per thread (16-byte accesses)
  template <class F>
  void ew4(const F& f, int64_t n) {
    if (n <= 0) return;
    const int64_t n4 = n / 4;
    int64_t blocks = (n4 + EW_THREADS - 1) / EW_THREADS;
    int64_t cap = (int64_t)sms * 32;
    if (blocks > cap) blocks = cap;
    ProfScope ps(PC_EW, st, 0.0, take_bytes());
    ew4_kernel<F><<<(unsigned)blocks, EW_THREADS, 0, st>>>(f, n4);
    post();
  }
  // C[M,N] (+)= A[M,K] B[K,N].  Shapes the tcgen05 tile covers (N % 128 == 0, K % 32 == 0, i.e.
  // hidden sizes that are multiples of 128) run on the tensor cores with 3xTF32 split
  // accumulation; everything else runs on the FP32 FFMA2 tile.
  void gemm_nn(const float* A, int64_t lda, const float* B, int64_t ldb, const float* Bt, int64_t ldbt, float* C,
               int64_t ldc, int64_t M, int N, int K, bool acc) {
    // gridDim.y carries the row tiles (<= 65535): very tall operands go in slabs
    constexpr int64_t SLAB = 65535LL * 128;
    for (int64_t m0 = 0; m0 < M; m0 += SLAB)
      gemm_nn_slab(A + m0 * lda, lda, B, ldb, Bt, ldbt, C + m0 * ldc, ldc, (M - m0 < SLAB) ? M - m0 : SLAB, N, K, acc);
  }
  void gemm_nn_slab(const float* A, int64_t lda, const float* B, int64_t ldb, const float* Bt, int64_t ldbt, float* C,
                    int64_t ldc, int64_t M, int N, int K, bool acc) {
    if (M <= 0) return;
    ProfScope ps(PC_STREAM_NN, st, 2.0 * M * N * K, 4.0 * M * (K + (acc ? 2.0 : 1.0) * N));
    if (use_tc && dgrad_res && N == dg::BN && K % (2 * dg::KC) == 0 && K > 128 && K <= dg::MAXCH * dg::KC) {
      // weights resident in shared memory, rows through tensor memory (the K = 3H data gradient of a DGM layer)
      if (first_use_on_device(3)) {
        note(cudaFuncSetAttribute(dg::dgrad_res_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dg::SMEM_BYTES));
        note(cudaFuncSetAttribute(dg::dgrad_res_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dg::SMEM_BYTES));
      }
      CUtensorMap tmA;
      if (dg::make_a_map(&tmA, A, lda, M, K)) {
        const int64_t ntiles = (M + dg::BM - 1) / dg::BM;
        int64_t npairs = sms / 2 > 0 ? sms / 2 : 1;
        if (npairs > ntiles) npairs = ntiles;
        if (acc) dg::dgrad_res_kernel<true><<<(unsigned)(2 * npairs), dg::NT, dg::SMEM_BYTES, st>>>(tmA, Bt, ldbt, hl_stride, C, ldc, M, K);
        else dg::dgrad_res_kernel<false><<<(unsigned)(2 * npairs), dg::NT, dg::SMEM_BYTES, st>>>(tmA, Bt, ldbt, hl_stride, C, ldc, M, K);
        post();
        return;
      }
    }
    if (use_tc && N % tc::BN == 0 && K % tc::KC == 0) {
      if (first_use_on_device(0)) {   // the opt-in is per function AND per device
        note(cudaFuncSetAttribute(tc::gemm_nn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
        note(cudaFuncSetAttribute(tc::gemm_nn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
      }
      dim3 grid(N / tc::BN, (unsigned)((M + tc::BM - 1) / tc::BM));
      if (acc) tc::gemm_nn_tc_kernel<true><<<grid, tc::NT, tc::SMEM_BYTES, st>>>(A, lda, Bt, ldbt, hl_stride, C, ldc, M, K);
      else tc::gemm_nn_tc_kernel<false><<<grid, tc::NT, tc::SMEM_BYTES, st>>>(A, lda, Bt, ldbt, hl_stride, C, ldc, M, K);
      post();
      return;
    }
    const int BN = (N % 128 == 0) ? 128 : (N % 64 == 0) ? 64 : 32;
    dim3 grid(N / BN, (unsigned)((M + GEMM_BM - 1) / GEMM_BM));
#define DGMK_NN(bn)                                                                                   \
  if (acc) gemm_nn_kernel<bn, true><<<grid, GEMM_NT, 0, st>>>(A, lda, B, ldb, C, ldc, M, K);          \
  else gemm_nn_kernel<bn, false><<<grid, GEMM_NT, 0, st>>>(A, lda, B, ldb, C, ldc, M, K);
    if (BN == 128) { DGMK_NN(128) } else if (BN == 64) { DGMK_NN(64) } else { DGMK_NN(32) }
#undef DGMK_NN
    post();
  }
  // ---- fused GEMM + element-wise stages (dgmk_lane_gemm.cuh): hidden size 128, channel sets
  // whose channel count divides 8 (value, ODE/FHN, heat) ------------------------------------
  bool lane_ok(int Hp, int cs) const {
    return use_tc && fuse && Hp == lg::KTOT && (cs == CS_V || cs == CS_D1O1 || cs == CS_HEAT);
  }
  // `units`: algorithmic HBM traffic of the launch in [M, 128] FP32 matrices (read + written)
  template <class EPI>
  int lane_gemm(const float* X, int64_t ldx, const float* Wt, int64_t ldw, int64_t M, int ngates, const EPI& epi, double units) {
    if (M <= 0) return 0;
    ProfScope ps(PC_LANE, st, 2.0 * M * lg::NU * lg::KTOT * ngates, units * M * lg::KTOT * 4.0);
    // opt-in shared memory size: per function and per device
    static std::atomic<unsigned long long> done_mask{0};   // one per EPI instantiation
    int dev = 0;
    note(cudaGetDevice(&dev));
    if (!((done_mask.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
      note(cudaFuncSetAttribute(lg::lane_gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, lg::SMEM_BYTES));
      done_mask.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    const int64_t ntiles = (M + lg::NR - 1) / lg::NR;
    int64_t grid, c0 = 0, c1 = 0;
    if (ngates == 1) {
      grid = ntiles < sms ? ntiles : sms;
    } else {   // three gates; the third (R: also forms s*R) costs ~40 % more epilogue time per tile
      // (measured with two MMA issuers, where the R gate's epilogue is what its CTAs wait for: 3.15 / 3.3 / 3.45 / 3.6 ->
      // lane class 48.8 / 46.5 / 47.6 / 48.1 ms per 2^20 heat rows; DGMK_GATE_SPLIT overrides for such sweeps)
      static const double split = [] { const char* e = getenv("DGMK_GATE_SPLIT"); double v = e ? atof(e) : 0.0; return v > 3.0 ? v : 3.3; }();
      c0 = (int64_t)(sms / split);
      if (c0 > ntiles) c0 = ntiles;
      if (c0 < 1) c0 = 1;
      c1 = c0;
      int64_t c2 = sms - 2 * c0;
      if (c2 > ntiles) c2 = ntiles;
      if (c2 < 1) c2 = 1;
      grid = c0 + c1 + c2;
    }
    lg::lane_gemm_kernel<EPI><<<(unsigned)grid, lg::NT, lg::SMEM_BYTES, st>>>(X, ldx, Wt, ldw, hl_stride, M, ngates, (int)c0, (int)c1, epi);
    post();
    return (int)grid;
  }
  // one DGM layer forward: [Z|G|R] GEMM + gate activations + s*R, then the H GEMM + activation +
  // state update.  Wb = packed [4*Hp, Hp] (gate, out unit) x in unit.
  template <class CS, int ACT>
  void dgm_fwd_fused(const XSrc& xs, const float* S, float* A4, const F4* ub, float* SR, float* Sn, const float* Wb, int Hp,
                     int64_t M) {
    if constexpr (CS::C == 1 || CS::C == 2 || CS::C == 4) {
      lg::DgmFwd1Epi<CS, ACT> e1; e1.xs = xs; e1.A4 = A4; e1.ub = ub; e1.S = S; e1.SR = SR;
      lane_gemm(S, Hp, Wb, Hp, M, 3, e1, 5.0);       // read s; write Z, G, R a-forms, s*R
      lg::DgmFwd2Epi<CS, ACT> e2; e2.xs = xs; e2.A4 = A4; e2.ub = ub; e2.S = S; e2.Sn = Sn;
      lane_gemm(SR, Hp, Wb + (int64_t)3 * Hp * Hp, Hp, M, 1, e2, 6.0);   // read s*R, Z, G, s; write H a-form, s'
    } else if (!err) err = "internal: fused path called with an unsupported channel set";
  }
  // (s*R)bar = abar_H W_h fused with the R-gate adjoint (DgmRev2Fn).  Wfh = packed [Hp in, Hp out]
  template <class CS, int ACT>
  void dgm_rev2_fused(const float* A4, const float* S, float* AB4, float* SBp, const float* Wfh, int Hp, int64_t M) {
    if constexpr (CS::C == 1 || CS::C == 2 || CS::C == 4) {
      lg::DgmRev2Epi<CS, ACT> e; e.A4 = A4; e.S = S; e.AB4 = AB4; e.SBp = SBp;
      lane_gemm(AB4 + 3 * Hp, 4 * (int64_t)Hp, Wfh, Hp, M, 1, e, 6.0);   // read abar_H, R a-form, s, s bar; write abar_R, s bar
    } else if (!err) err = "internal: fused path called with an unsupported channel set";
  }
  // DgmRev1Fn + grad[U | b] of the Z, G, H gates (hidden size 128; rows = collocation points)
  template <class CS, class F>
  void dgm_rev1_e(const F& f, const XSrc& xs, int64_t rows, float* outE, float* part, int64_t part_n) {
    const int64_t n = rows * 128;
    if (n <= 0) return;
    if constexpr (!(CS::C == 1 || CS::C == 2 || CS::C == 4)) { if (!err) err = "internal: fused path called with an unsupported channel set"; return; } else {
    // units per thread: 4 (value-only and ODE / FHN rows), 2 (heat: four channels per unit)
    constexpr int V = (CS::C <= 2) ? 4 : 2;
    constexpr int MB = (CS::C == 1) ? 3 : 2;   // resident blocks per SM (80 / 126 registers)
    const int64_t nv = rows * (128 / V);
    int64_t blocks = (nv + EW_THREADS - 1) / EW_THREADS;
    int64_t cap = (int64_t)sms * MB;            // one resident wave
    if (cap > part_n / (9 * 128)) cap = part_n / (9 * 128);
    if (blocks > cap) blocks = cap;
    if (blocks < 1) { if (!err) err = "internal: partial buffer too small"; return; }
    {
      ProfScope ps(PC_EW, st, 0.0, 9.0 * rows * CS::C * 128 * 4.0);   // read s'bar, Z, G, H, s; write abar_Z, abar_G, abar_H, s bar
      rev1_ev_kernel<F, CS, V, MB><<<(unsigned)blocks, EW_THREADS, 0, st>>>(f, xs, nv, part);
      post();
    }
    reduce_gate_e(part, (int)blocks, 3, 0, 1, 3, outE, 4 * 128);
    }
  }
  void reduce_gate_e(const float* part, int nparts, int ngates, int gm0, int gm1, int gm2, float* out, int64_t ldo) {
    ProfScope ps(PC_OTHER, st, 0.0, 4.0 * nparts * ngates * 384);
    reduce_gate_e_kernel<<<(unsigned)((ngates * 384 + 31) / 32), dim3(32, nparts >= 512 ? 32 : 8), 0, st>>>(part, nparts, ngates, gm0, gm1, gm2, out, ldo);
    post();
  }
  // MLP hidden layer forward: GEMM + bias + activation
  template <class CS, int ACT>
  void mlp_fwd_fused(const float* Yp, float* G, const F4* ub, float* Yn, const float* Wb, int Hp, int64_t M) {
    if constexpr (CS::C == 1 || CS::C == 2 || CS::C == 4) {
      lg::MlpActEpi<CS, ACT> e; e.G = G; e.ub = ub; e.Yn = Yn;
      lane_gemm(Yp, Hp, Wb, Hp, M, 1, e, 3.0);
    } else if (!err) err = "internal: fused path called with an unsupported channel set";
  }
  // MLP reverse, layers below the top one: ABout = act_adj(ABin Wt^T, a-form G of the layer below) in one launch
  template <class CS, int ACT>
  void mlp_rev_fused(const float* ABin, const float* Wt, const float* G, float* ABout, int Hp, int64_t M) {
    if constexpr (CS::C == 1 || CS::C == 2 || CS::C == 4) {
      lg::MlpRevEpi<CS, ACT> e; e.G = G; e.AB = ABout;
      lane_gemm(ABin, Hp, Wt, Hp, M, 1, e, 3.0);   // read abar_l, the a-form below; write abar_{l-1}
    } else if (!err) err = "internal: fused path called with an unsupported channel set";
  }
  // MLP reverse, bottom layer: ABout = act_adj(ABin Wt^T, input-layer a-form) in one launch (lane_store + InputRevFn)
  template <class CS, int ACT>
  void input_rev_fused(const float* ABin, const float* Wt, const F4* inb, const float* S0, float* ABout, int Hp, int64_t M) {
    if constexpr (CS::C == 1 || CS::C == 2 || CS::C == 4) {
      lg::InputRevEpi<CS, ACT> e; e.inb = inb; e.S0 = S0; e.AB = ABout;
      lane_gemm(ABin, Hp, Wt, Hp, M, 1, e, 2.0 + 1.0 / CS::C);   // read abar_0, the value rows of s0; write abar_in
    } else if (!err) err = "internal: fused path called with an unsupported channel set";
  }
  // C[M, Hp] = X[M, Hp] Wt[Hp, Hp]^T on the lane kernel (MLP data gradient)
  void lane_store(const float* X, int64_t ldx, const float* Wt, float* C, int64_t ldc, int Hp, int64_t M) {
    lg::StoreEpi<false> e; e.C = C; e.ldc = ldc;
    lane_gemm(X, ldx, Wt, Hp, M, 1, e, 2.0);
  }
  // out[(i / cols) * ldo + i % cols] += sum_p part[p][i]   (cols = 0: out[i])
  void reduce(const float* part, int nparts, int64_t n, float* out, int cols = 0, int64_t ldo = 0, bool overwrite = false) {
    if (cols <= 0) { cols = (int)n; ldo = n; }
    // slices: enough to keep every chain short, few enough that small partial counts are not split to nothing
    const int S = nparts >= 512 ? 32 : (nparts >= 128 ? 16 : (nparts >= 16 ? 8 : 1));
    ProfScope ps(PC_OTHER, st, 0.0, 4.0 * nparts * (double)n);
    reduce_partials_kernel<<<(unsigned)((n + 31) / 32), dim3(32, S), 0, st>>>(part, nparts, n, out, cols, ldo, overwrite);
    post();
  }
  // out[N,Kd] += A^T S ; if E: outE[e * ldoE + n] += sum_m A[m,n] E[m,e]  (fused in the same pass)
  void gemm_tn_acc(const float* A, int64_t lda, const float* S, int64_t lds, float* out, int N, int Kd, int64_t M,
                   const float* E, float* outE, int64_t ldoE, float* part, int64_t part_n) {
    if (M <= 0) return;
    const int64_t tile = (int64_t)N * Kd;
    const int64_t per_split = tile + (E ? 4 * (int64_t)N : 0);
    int64_t max_splits = part_n / per_split;
    if (max_splits > 2048) max_splits = 2048;
    if (max_splits < 1) { if (!err) err = "internal: partial buffer too small"; return; }
    // one FP32 partial per ~1024 rows (as many as the partial buffer holds): a split's accumulators are plain
    // FP32 chains over its rows, and 8192-row chains put grad b of MLP(1,1,32) 3e-5 .. 5e-5 from FP64 at 2^20
    // rows (the reference's own FP32: 5e-7); the partials are added in FP64
    const int BN = (Kd % 128 == 0) ? 128 : (Kd % 64 == 0) ? 64 : 32;
    int64_t by_rows = (M + 1023) / 1024;
    int64_t splits = by_rows;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int64_t rps = ((M + splits - 1) / splits + 31) / 32 * 32;
    splits = (M + rps - 1) / rps;
    float* PE = E ? part + splits * tile : nullptr;
    // warp-specialised tcgen05 kernel (A^T through tensor memory): one CTA per SM, one wave
    const bool ws_ok = use_tc && fuse && N % tc::BM == 0 && Kd % tc::BN == 0 && lds == 128 && (lda == 128 || lda == 512);
    if (ws_ok) {
      // CTAs: one wave; each walks `nseg` segments of seg_rows rows and emits one partial per
      // segment, so the FP32 chains stay as short as with the many-splits streaming tile
      const int tiles_ws = (Kd / 128) * (N / 128);
      const int64_t per_seg = tile + 8 * (int64_t)N;    // W partial + two E partials
      int64_t max_seg = part_n / per_seg;
      if (max_seg > 512) max_seg = 512;
      int64_t ctas = sms / tiles_ws;
      if (ctas > by_rows) ctas = by_rows;
      if (ctas > max_seg) ctas = max_seg;
      if (ctas < 1) ctas = 1;
      int64_t nseg = ((M + ctas - 1) / ctas + 8191) / 8192;   // segments of <= 8192 rows
      if (nseg * ctas > max_seg) nseg = max_seg / ctas;
      if (nseg < 1) nseg = 1;
      const int64_t seg_rows = ((M + ctas * nseg - 1) / (ctas * nseg) + 31) / 32 * 32;
      ctas = (M + seg_rows * nseg - 1) / (seg_rows * nseg);
      const int64_t nparts = (M + seg_rows - 1) / seg_rows;   // segments that contain rows (the rest is never written)
      float* PEw = E ? part + ctas * nseg * tile : nullptr;
      dim3 gws(Kd / 128, N / 128, (unsigned)ctas);
      if (first_use_on_device(2)) {
        note(cudaFuncSetAttribute(wg::wgrad_ws_kernel<512, 128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::SMEM_BYTES));
        note(cudaFuncSetAttribute(wg::wgrad_ws_kernel<128, 128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::SMEM_BYTES));
        note(cudaFuncSetAttribute(wg::wgrad_ws_kernel<512, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::SMEM_BYTES_SEP));
        note(cudaFuncSetAttribute(wg::wgrad_ws_kernel<128, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::SMEM_BYTES_SEP));
      }
      ProfScope ps(PC_WGRAD, st, 2.0 * M * N * Kd, 4.0 * M * (N + Kd + (E ? 4.0 : 0.0)));
      // without the A^T E side product: the variant with its own MMA-issuer warpgroup
      if (!E) {
        if (lda == 512) wg::wgrad_ws_kernel<512, 128, true><<<gws, wg::NT_SEP, wg::SMEM_BYTES_SEP, st>>>(A, S, nullptr, part, nullptr, N, Kd, M, seg_rows, (int)nseg);
        else wg::wgrad_ws_kernel<128, 128, true><<<gws, wg::NT_SEP, wg::SMEM_BYTES_SEP, st>>>(A, S, nullptr, part, nullptr, N, Kd, M, seg_rows, (int)nseg);
      } else if (lda == 512) wg::wgrad_ws_kernel<512, 128, false><<<gws, wg::NT, wg::SMEM_BYTES, st>>>(A, S, E, part, PEw, N, Kd, M, seg_rows, (int)nseg);
      else wg::wgrad_ws_kernel<128, 128, false><<<gws, wg::NT, wg::SMEM_BYTES, st>>>(A, S, E, part, PEw, N, Kd, M, seg_rows, (int)nseg);
      post();
      reduce(part, (int)nparts, tile, out);
      if (E) {
        reduce(PEw, (int)(2 * nparts), 4 * (int64_t)N, outE, N, ldoE);
      }
      return;
    }
    dim3 grid(Kd / BN, (N + GEMM_BM - 1) / GEMM_BM, (unsigned)splits);
    if (use_tc && E && N % tc::BM == 0 && Kd % tc::BN == 0) {
      if (first_use_on_device(1)) {
        note(cudaFuncSetAttribute(tctn::gemm_tn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tctn::TN_SMEM_BYTES));
      }
      tctn::gemm_tn_tc_kernel<<<grid, tctn::NT, tctn::TN_SMEM_BYTES, st>>>(A, lda, S, lds, E, part, PE, N, Kd, M, rps);
    } else if (BN == 128) gemm_tn_kernel<128><<<grid, GEMM_NT, 0, st>>>(A, lda, S, lds, part, N, Kd, M, rps, E, PE);
    else if (BN == 64) gemm_tn_kernel<64><<<grid, GEMM_NT, 0, st>>>(A, lda, S, lds, part, N, Kd, M, rps, E, PE);
    else gemm_tn_kernel<32><<<grid, GEMM_NT, 0, st>>>(A, lda, S, lds, part, N, Kd, M, rps, E, PE);
    post();
    reduce(part, (int)splits, tile, out);
    if (E) {
      reduce(PE, (int)splits, 4 * (int64_t)N, outE, N, ldoE);
    }
  }
  // ldo > 0: out[e * ldo + n] (strided rows) instead of the packed [NE][N]
  void wcolsum_acc(const float* Mat, int64_t ldm, int N, const float* Wt, int64_t M, float* out, float* part,
                   int64_t part_n, int64_t ldo = 0) {
    if (M <= 0) return;
    const int NE = Wt ? 4 : 1;
    int64_t max_blk = part_n / ((int64_t)NE * N);
    if (max_blk > 2048) max_blk = 2048;   // enough loads in flight to stream at HBM speed (second stage: nblk partials)
    if (max_blk < 1) { if (!err) err = "internal: partial buffer too small"; return; }
    const int ctiles = (N + 127) / 128;
    int64_t want = ((int64_t)sms * 32 + ctiles - 1) / ctiles;
    int64_t by_rows = (M + 255) / 256;
    int64_t nblk = want < by_rows ? want : by_rows;
    if (nblk > max_blk) nblk = max_blk;
    if (nblk < 1) nblk = 1;
    int64_t rpb = (M + nblk - 1) / nblk;
    nblk = (M + rpb - 1) / rpb;
    dim3 grid(ctiles, (unsigned)nblk);
    {
      ProfScope ps(PC_OTHER, st, 0.0, 4.0 * M * (N + (Wt ? 4.0 : 0.0)));
      if (Wt) wcolsum_kernel<true><<<grid, 128, 0, st>>>(Mat, ldm, N, Wt, M, rpb, part);
      else wcolsum_kernel<false><<<grid, 128, 0, st>>>(Mat, ldm, N, Wt, M, rpb, part);
      post();
    }
    if (ldo > 0) reduce(part, (int)nblk, (int64_t)NE * N, out, N, ldo);
    else reduce(part, (int)nblk, (int64_t)NE * N, out);
  }
  void rowdot(const float* S, int64_t lds, const float* W, const float* b, float* U, int64_t M, int Hp, int o, int C) {
    if (M <= 0) return;
    int64_t blocks = (M + 7) / 8;
    int64_t cap = (int64_t)sms * 16;
    if (blocks > cap) blocks = cap;
    ProfScope ps(PC_OTHER, st, 0.0, 4.0 * M * (Hp + 4.0));
    rowdot_kernel<<<(unsigned)blocks, 256, 0, st>>>(S, lds, W, b, U, M, Hp, o, C);
    post();
  }
  // ---- resident-tile step (dgmk_tile.cuh): hidden sizes <= 64, the whole step in one persistent kernel --------
  static constexpr bool kHasTile = true;
  // Plans the shared-memory layout, launches the kernel on c.Wp (packed) and adds the per-CTA partials into c.Gp.
  // false = shape not covered (the caller runs the layer-wise path).
  // shared-memory budget left for the tile region after the fixed parts (scratch, staged weights, accumulators)
  int64_t tile_fixed(const Ctx& c, tk::TileParams& prm, int64_t smem_limit = tk::SMEM_MAX) {
    prm.n = c.n; prm.pl = c.pl; prm.Wp = c.Wp; prm.slots = c.part;
    prm.w_floats = (uint32_t)((c.pl.w_total + 3) / 4 * 4);
    prm.g_floats = (uint32_t)c.pl.g_total;
    int64_t budget = smem_limit - (int64_t)tk::SCRATCH_TOTAL_FLOATS * 4;
    prm.w_smem = (int64_t)prm.w_floats * 4 <= 64 * 1024;
    if (prm.w_smem) budget -= (int64_t)prm.w_floats * 4;
    prm.g_smem = (int64_t)prm.g_floats * 4 <= 32 * 1024;
    if (prm.g_smem) budget -= (int64_t)prm.g_floats * 4;
    prm.lp_floats = prm.coord_floats = prm.ip_floats = 0; prm.J = 0;
    memset(&prm.heat, 0, sizeof(prm.heat)); memset(&prm.ode, 0, sizeof(prm.ode)); memset(&prm.fred, 0, sizeof(prm.fred));
    prm.prof = g_tile_prof; prm.prof_n = g_tile_prof_n;
    return budget;
  }
  // grid / slot bookkeeping, launch, second-stage reduction of the per-CTA partials into c.Gp
  bool tile_launch(Ctx& c, int prob, tk::TileParams& prm, double falg, double balg) {
    const int64_t ntiles = (prm.B + prm.P - 1) / prm.P;
    const size_t smem_need = (size_t)tk::SCRATCH_TOTAL_FLOATS * 4 + (prm.w_smem ? (size_t)prm.w_floats * 4 : 0) + (prm.g_smem ? (size_t)prm.g_floats * 4 : 0) +
                             ((size_t)prm.lp_floats + prm.coord_floats + prm.ip_floats) * 4 + prm.tile_bytes;
    // two CTAs per SM when the plan fits half an SM's shared memory (the kernels that allow it are compiled for it)
    const bool two = prob != tk::PROB_HEAT && !(DGMK_TILE_DGM_ONE_CTA && c.n.is_dgm()) && smem_need <= (size_t)tk::SMEM_HALF;
    const int64_t ctas = (int64_t)sms * (two ? 2 : 1);
    int64_t grid = ntiles < ctas ? ntiles : ctas;
    const int64_t slots_avail = c.part_n / prm.g_floats;
    if (slots_avail < 1) return false;
    if (grid > slots_avail) grid = slots_avail;   // (wide networks: fewer partial-gradient slots than SMs; CTAs walk more tiles)
    int64_t nseg = ((ntiles + grid - 1) / grid + g_tile_flush - 1) / g_tile_flush;
    if (nseg * grid > slots_avail) nseg = slots_avail / grid;
    prm.nslots_per_cta = (int32_t)nseg;
    prm.flush_tiles = g_tile_flush;
    const size_t smem = smem_need;
    note(cudaMemsetAsync(c.part, 0, (size_t)grid * nseg * prm.g_floats * 4, st));
    {
      ProfScope ps(PC_TILE, st, falg, balg);
      note((cudaError_t)tk::launch(prob, prm, (int)grid, smem, st));
      ++g_launches;
    }
    // The packed gradient IS the sum of the partials (the caller has not zeroed c.Gp).  Few partials (small batches,
    // where every launch counts): sum, scatter into the caller's gradient and write the loss in one launch.
    const int64_t nparts = grid * nseg;
    unpacked = false;
    if (nparts <= 64 && (out_grad || out_loss)) {
      const int64_t np = num_params(c.n);
      ProfScope ps(PC_OTHER, st, 0.0, 4.0 * nparts * (double)np);
      tile_unpack_kernel<<<(unsigned)((np + 1 + 127) / 128), 128, 0, st>>>(c.grad, c.part, (int)nparts, prm.g_floats, c.pl.g_acc, out_grad, out_loss, np);
      post();
      unpacked = true;
    } else {
      reduce(c.part, (int)nparts, prm.g_floats, c.Gp, 0, 0, true);
    }
    return !err;
  }
  // where a tile step may deliver the flat gradient and the loss directly (set by the C API before tile_step*)
  float* out_grad = nullptr; float* out_loss = nullptr; bool unpacked = false;
  void set_unpack_target(float* grad, float* loss) { out_grad = grad; out_loss = loss; }
  static int64_t r4(int64_t v) { return (v + 3) / 4 * 4; }
  // Heat / ODE / FitzHugh-Nagumo.  false = shape not covered or the layer-wise path is faster (the caller runs it).
  bool tile_step(Ctx& c, int cls, const HeatArgs* ha, const OdeArgs* oa, int64_t B) {
    if (!g_tile || c.n.Hp > TILE_MAX_HP || B <= 0) return false;
    // Hidden sizes above 64 (the headline 128) fit the tile step for small batches only (4 heat points per tile, weights
    // and gradient accumulators through L2).  Measured at the reference's own 64 rows, DGM(2,1,128,3) under the CUDA-graph
    // driver: 0.99 ms per iteration against 0.85 ms for the ~115 tcgen05 launches of the layer-wise path -- every GEMM
    // stage waits on weight rows coming from L2 (ldg chains, two k-steps in flight).  So it is OFF by default and runs
    // only when forced (dgmk_set_tile_engine(2): the parity tests cover it); staging weight panels through shared
    // memory is the missing piece.
    if (c.n.Hp > TILE_WIDE_HP && (g_tile != 2 || B > TILE_WIDE_ROWS)) return false;
    tk::TileParams prm;
    // per-point extras: loss contribution per loss row + the staged coordinates of the larger pass
    auto coord_fl = [&](int64_t P) { return r4((cls == DGMK_WS_HEAT ? 6 : 1) * P); };
    auto need = [&](int64_t P) { return (int64_t)chunk_region_bytes(c.n, cls, P, 0, true) + r4(loss_points(cls, P, 0)) * 4 + coord_fl(P) * 4; };
    const int64_t Pcap = B < 128 ? B : 128;
    auto plan = [&](int64_t limit) {
      const int64_t budget = tile_fixed(c, prm, limit);
      int64_t P = 0;
      for (int64_t q = 1; q <= Pcap; ++q) { if (need(q) <= budget) P = q; else break; }
      return P;
    };
    // Two CTAs per SM hide each other's stage barriers (simple_ode MLP(1,1,32): 4.5e8 -> 5.0e8 rows/s) as long as a
    // half-SM tile still has >= 128 interior rows; smaller tiles lose more than the overlap gains (measured: heat
    // DGM(2,1,32,1) 15.7 -> 24.2 ms with 11-point tiles), and the heat kernels are compiled for one CTA per SM.
    int64_t P = (cls == DGMK_WS_HEAT || (DGMK_TILE_DGM_ONE_CTA && c.n.is_dgm())) ? 0 : plan(tk::SMEM_HALF);
    if (P * 2 < 128 && P < B) P = plan(tk::SMEM_MAX);
    prm.B = B;
    if (P < 2 && P < B) return false;
    // Measured crossover (tools/tile_sweep.py): with the whole stash in shared memory the tiles of wide / deep
    // networks get small (7 heat points at hidden size 64, 3 layers) and past a few thousand rows the layer-wise
    // path, which streams large GEMMs through HBM, is faster; at hidden size 32 the tile step wins or ties at
    // every batch size.  g_tile == 2 forces the tile step.
    const int64_t tile_rows = P * (cls == DGMK_WS_HEAT ? 4 : 2);
    if (g_tile == 1 && c.n.Hp > 32 && B > 4096 && !(tile_rows >= 48 && B <= 16384)) return false;
    // Small batches (the reference's own 32 .. 256 rows): a stage costs ~1000 cycles however few rows it has, so one
    // full tile on one SM is the slowest possible plan -- spread the batch over the SMs in tiles of >= 4 points
    // (64 rows: 16 CTAs of 4 points instead of 1 CTA of 64)
    {
      const int64_t per_cta = (B + sms - 1) / sms;
      const int64_t Ps = per_cta < 4 ? 4 : per_cta;
      if (Ps < P) P = Ps;
    }
    // spread the points evenly over the tiles (same tile count, smaller ragged tail)
    const int64_t nt0 = (B + P - 1) / P;
    P = (B + nt0 - 1) / nt0;
    prm.P = (int32_t)P;
    prm.lp_floats = (uint32_t)r4(loss_points(cls, P, 0));
    prm.coord_floats = (uint32_t)coord_fl(P);
    prm.tile_bytes = (uint32_t)chunk_region_bytes(c.n, cls, P, 0, true);
    if (ha) prm.heat = *ha;
    if (oa) prm.ode = *oa;
    // algorithmic flops: 3 M (L c H^2 + 2 H o) per row; algorithmic bytes: the point coordinates
    const double Mrows = cls == DGMK_WS_HEAT ? 7.0 : 3.0, cc = c.n.is_dgm() ? 8.0 : 2.0;
    const double falg = 3.0 * Mrows * (c.n.L * cc * c.n.H * c.n.H + 2.0 * c.n.H * c.n.o) * (double)B;
    const double balg = (cls == DGMK_WS_HEAT ? 40.0 : (cls == DGMK_WS_FHN ? 16.0 : 12.0)) * (double)B;
    return tile_launch(c, cls == DGMK_WS_HEAT ? tk::PROB_HEAT : tk::PROB_ODE, prm, falg, balg);
  }
  // Fredholm: a tile is a block of P points with all their k nodes, the node rows taken J nodes at a time
  bool tile_step_fredholm(Ctx& c, const FredArgs& fa) {
    if (!g_tile || c.n.Hp > TILE_WIDE_HP || fa.B <= 0) return false;
    tk::TileParams prm;
    const int64_t budget = tile_fixed(c, prm);
    prm.B = fa.B;
    auto need = [&](int64_t P, int64_t J) {
      return (int64_t)pass_bytes(c.n, P, CS_V) + (int64_t)pass_bytes(c.n, P * J, CS_V) + (int64_t)rev_bytes(c.n, P * J, true) +
             (r4(P) * 2 + r4(P * J)) * 4;
    };
    // points per block: at most 8, and few enough that a small batch still spreads over the SMs -- the reference's own
    // step (32 points x 50 nodes) is 32 one-point blocks whose 50 node rows fit ONE sub-tile (no second node pass)
    // instead of four 8-point blocks of three sub-tiles each
    int64_t P = (fa.B + sms - 1) / sms, J = 0;
    if (P > 8) P = 8;
    if (P < 1) P = 1;
    for (int64_t q = 1; q <= fa.k; ++q) { if (need(P, q) <= budget) J = q; else break; }
    if (J < 1) return false;
    // measured (tools/tile_prof.py fredholm): 8-point blocks with ~17-node sub-tiles and the second node pass lose to
    // the layer-wise path's large GEMMs once the step has more than ~64 K node evaluations (k = 1024, B = 2^14:
    // 66 ms against 46 ms); below that the single launch wins (B = 32, k = 50: 0.30 ms against 1.8 ms)
    if (g_tile == 1 && fa.B * (int64_t)(fa.k + 1) > 65536) return false;
    const int64_t nsub = (fa.k + J - 1) / J;   // even sub-tiles
    J = (fa.k + nsub - 1) / nsub;
    prm.P = (int32_t)P; prm.J = (int32_t)J;
    prm.lp_floats = (uint32_t)r4(P); prm.ip_floats = (uint32_t)r4(P); prm.coord_floats = (uint32_t)r4(P * J);
    prm.tile_bytes = (uint32_t)(pass_bytes(c.n, P, CS_V) + pass_bytes(c.n, P * J, CS_V) + rev_bytes(c.n, P * J, true));
    prm.fred = fa;
    const double falg = 3.0 * (fa.k + 1.0) * (c.n.L * (c.n.is_dgm() ? 8.0 : 2.0) * c.n.H * c.n.H + 2.0 * c.n.H * c.n.o) * (double)fa.B;
    return tile_launch(c, tk::PROB_FRED, prm, falg, 4.0 * (fa.k + 1.0) * (double)fa.B);
  }
  void zero(void* p, size_t bytes) { note(cudaMemsetAsync(p, 0, bytes, st)); }
  void copy(void* dst, const void* src, size_t bytes) { note(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st)); }
  bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
  }
  // consumes the sticky-less CUDA error state so that a failed launch here does not leak into the caller's
  // (PyTorch's) own launch checks or into later dgmk calls
  const char* error() { note(cudaGetLastError()); return err; }
};

}  // namespace dgmk

DGMK_DEFINE_C_API(dgmk::CudaBackend, "cuda-sm100a")

// ---- diagnostics (not part of the reference-facing surface) -------------------------
namespace dgmk {
// dependent-chain FFMA: 16 independent accumulators per thread, operands in registers
__global__ void __launch_bounds__(256) ffma_probe_kernel(float* out, const float* in, int iters) {
  float acc[16];
  float b = in[threadIdx.x % 7], c = in[threadIdx.x % 5 + 1];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = in[i];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], b, c);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace dgmk

extern "C" {
unsigned long long dgmk_launch_count(void) { return dgmk::g_launches; }
// Per-kernel-class timing.  dgmk_profile(1) clears the counters and starts recording CUDA events
// around every launch; dgmk_profile(0) stops.  dgmk_profile_read(cls, ...) synchronises the
// recorded events and returns the class's summed duration [ms], launches, algorithmic flops and
// bytes.  Classes: 5 resident-tile step kernel (hidden sizes <= 64), 0 weight gradient (wgrad_ws), 1 fused units-on-lanes GEMM + element-wise kernels,
// 2 streaming tcgen05 / FFMA GEMM tiles, 3 stand-alone element-wise kernels, 4 reductions / output layer.
void dgmk_profile(int on) {
  using namespace dgmk;
  if (on) {
    for (int c = 0; c < PC_COUNT; ++c) {
      for (cudaEvent_t e : g_prof[c].ev) g_prof_pool.push_back(e);
      g_prof[c] = ProfClass();
    }
  }
  g_prof_on = on != 0;
}
int dgmk_profile_read(int cls, double* ms, long long* launches, double* flops, double* bytes) {
  using namespace dgmk;
  if (cls < 0 || cls >= PC_COUNT) return DGMK_EINVAL;
  ProfClass& p = g_prof[cls];
  double total = 0.0;
  for (size_t i = 0; i + 1 < p.ev.size(); i += 2) {
    if (cudaEventSynchronize(p.ev[i + 1]) != cudaSuccess) return DGMK_ECUDA;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, p.ev[i], p.ev[i + 1]) != cudaSuccess) return DGMK_ECUDA;
    total += t;
  }
  if (ms) *ms = total;
  if (launches) *launches = p.launches;
  if (flops) *flops = p.flops;
  if (bytes) *bytes = p.bytes;
  return 0;
}
// 0 = FP32 FFMA2 tiles only; 1 = tcgen05 3xTF32, fused GEMM + element-wise kernels where the shape
// allows (default); 2 = tcgen05 3xTF32 streaming tiles + separate element-wise kernels; 3 = like 1, but the K = 3H data
// gradient on the round-1 streaming tile instead of the weight-resident kernel (A/B measurements)
void dgmk_set_gemm_engine(int engine) {
  dgmk::g_use_tc = engine != 0; dgmk::g_fuse = engine == 1 || engine == 3; dgmk::g_dgrad_res = engine == 1;
}
// 1 (default) = hidden sizes <= 64 run the resident-tile step (one persistent kernel per step, dgmk_tile.cuh);
// 0 = the layer-wise path for every hidden size (A/B measurements, tests of the layer-wise path at small sizes);
// 2 = the resident-tile step wherever a tile fits, whatever the dispatch rule says
void dgmk_set_tile_engine(int on) { dgmk::g_tile = on; }
// diagnostic: CTA 0 of the following resident-tile launches writes (clock64, stage kind) pairs -- one per stage, at most
// n -- into buf (device memory, 2 * n int64); kinds: 0 start, 1 ew, 2 ew4, 3 gemm_nn, 4 column sums, 5 gemm_tn, 6 A^T E,
// 7 rowdot.  buf = NULL switches it off.
// tiles per FP32 accumulation segment of the resident-tile step (default 256; tests lower it to exercise the segment
// hand-over on small batches)
void dgmk_set_tile_flush(int tiles) { dgmk::g_tile_flush = tiles > 0 ? tiles : dgmk::tk::FLUSH_TILES; }
void dgmk_tile_profile(long long* buf, int n) { dgmk::g_tile_prof = buf; dgmk::g_tile_prof_n = buf ? n : 0; }
// same tcgen05 tile the pipeline launches: C[M,N] = A[M,K] Bt[N,K]^T, lda = ldc = ld
int dgmk_gemm_tc_probe(const float* A, const float* Bt, float* C, int64_t M, int N, int K, int64_t ld, void* stream) {
  if (N % 128 || K % 32) return DGMK_EINVAL;
  dgmk::CudaBackend bk(stream);
  bk.use_tc = true;
  bk.hl_stride = (int64_t)N * K;  // Bt = [plain | tf32-hi | tf32-lo], each N*K floats
  bk.gemm_nn(A, ld, nullptr, 0, Bt, K, C, ld, M, N, K, false);
  return bk.error() ? DGMK_ECUDA : 0;
}
// FP32 FFMA peak probe: launches blocks x 256 threads, each doing 64*iters FFMAs.
// `in` >= 32 floats, `out` >= blocks*256 floats (device).  flops = 2*64*iters*256*blocks.
int dgmk_ffma_probe(const float* in, float* out, int blocks, int iters, void* stream) {
  dgmk::ffma_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, in, iters);
  return cudaPeekAtLastError() == cudaSuccess ? 0 : DGMK_ECUDA;
}
// The dominant GEMM tile stand-alone (same kernel the pipeline launches) for the live
// roofline measurement in bench.py: C[M,N] = A[M,K] B[K,N], lda = ldc = ld.
int dgmk_gemm_probe(const float* A, const float* B, float* C, int64_t M, int N, int K, int64_t ld, void* stream) {
  if (N % 32 || K % 16) return DGMK_EINVAL;
  dgmk::CudaBackend bk(stream);
  bk.use_tc = false;  // the FFMA2 tile; B is [K,N]
  bk.gemm_nn(A, ld, B, N, nullptr, 0, C, ld, M, N, K, false);
  return bk.error() ? DGMK_ECUDA : 0;
}
}
