// Instantiations of the resident-tile step kernels (dgmk_tile.cuh) and their launcher.  A separate translation
// unit so that the two halves of libdgmk.so compile in parallel; linked into the same library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -c dgmk_tile.cu
#include <atomic>
#define DGMK_TILE_TU 1   // dgmk_math.h: functor buffers are shared memory in this translation unit
#include "dgmk_tile.cuh"

namespace dgmk {
namespace tk {

template <int PROB, class BK>
static cudaError_t launch_one(const TileParams& prm, int grid, size_t smem, cudaStream_t st) {
  static std::atomic<unsigned long long> done_mask{0};   // the shared-memory opt-in is per function AND per device
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (!((done_mask.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
    e = cudaFuncSetAttribute(tile_step_kernel<PROB, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    if (e != cudaSuccess) return e;
    done_mask.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
  tile_step_kernel<PROB, BK><<<grid, NT, smem, st>>>(prm);
  return cudaPeekAtLastError();
}

template <int PROB>
static cudaError_t launch_prob(const TileParams& prm, int grid, size_t smem, cudaStream_t st) {
  constexpr int CSM = (PROB == PROB_HEAT) ? ((1 << CS_HEAT) | (1 << CS_V)) : (PROB == PROB_ODE ? ((1 << CS_D1O1) | (1 << CS_V)) : (1 << CS_V));
  const NetDims& n = prm.n;
  if (n.kind == KIND_MLP) {
    switch (n.act) {
      case ACT_RELU: return launch_one<PROB, TileBackend<CSM, 1 << ACT_RELU, true, false>>(prm, grid, smem, st);
      case ACT_SIGMOID: return launch_one<PROB, TileBackend<CSM, 1 << ACT_SIGMOID, true, false>>(prm, grid, smem, st);
      case ACT_TANH: return launch_one<PROB, TileBackend<CSM, 1 << ACT_TANH, true, false>>(prm, grid, smem, st);
      default: return launch_one<PROB, TileBackend<CSM, 1 << ACT_LEAKY, true, false>>(prm, grid, smem, st);
    }
  }
  if (n.kind == KIND_DGM_LINEAR) return launch_one<PROB, TileBackend<CSM, 1 << ACT_TANH, false, true>>(prm, grid, smem, st);
  // neural_networks.DGM: ReLU gates; the input layer is ReLU (as shipped) or tanh (func != "relu")
  if (n.act == ACT_TANH) return launch_one<PROB, TileBackend<CSM, (1 << ACT_RELU) | (1 << ACT_TANH), false, true>>(prm, grid, smem, st);
  return launch_one<PROB, TileBackend<CSM, 1 << ACT_RELU, false, true>>(prm, grid, smem, st);
}

// returns a cudaError_t as int
int launch(int prob, const TileParams& prm, int grid, size_t smem, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (prob == PROB_HEAT) return (int)launch_prob<PROB_HEAT>(prm, grid, smem, st);
  if (prob == PROB_ODE) return (int)launch_prob<PROB_ODE>(prm, grid, smem, st);
  return (int)launch_prob<PROB_FRED>(prm, grid, smem, st);
}

}  // namespace tk
}  // namespace dgmk
