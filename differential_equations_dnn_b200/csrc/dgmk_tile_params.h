// Launch parameters and shared-memory budget of the resident-tile step kernels (dgmk_tile.cuh): the part the host
// side (CudaBackend::tile_step in dgmk_cuda.cu) and the kernels' translation unit (dgmk_tile.cu) share.
#pragma once
#include <stdint.h>
#include "dgmk_steps.h"

namespace dgmk {
namespace tk {

constexpr int NT = 256;                 // threads per CTA, one CTA per SM (measured: 512 threads 19.3 ms, 1024 25.4 ms, 256 16.2 ms
                                        // per 2^20 heat rows at hidden size 32 -- fewer idle lanes at the stage barriers, no spills)
constexpr int SCRATCH_FLOATS = 2048;    // cross-group reduction scratch (8 KB)
#ifndef DGMK_TILE_SCRATCH2
#define DGMK_TILE_SCRATCH2 0
#endif
// second scratch: > 0 lets the A^T E column sums share a stage with the weight-gradient tiles (1536 floats cover every
// shape).  Measured (profiles/r02_notes.md): the stage it saves is worth less than the ~6 % of tile points the 6 KB cost.
constexpr int SCRATCH2_FLOATS = DGMK_TILE_SCRATCH2;
constexpr int SCRATCH_TOTAL_FLOATS = SCRATCH_FLOATS + SCRATCH2_FLOATS;
constexpr int SMEM_MAX = 232448;        // 227 KB opt-in limit per CTA on sm_100
constexpr int SMEM_HALF = 115712;       // two CTAs per SM: (228 KB - 2 x 1 KB reserved) / 2
constexpr int FLUSH_TILES = 256;        // tiles per FP32 accumulation segment

// 1: the DGM kernels of the ODE / Fredholm classes get one CTA per SM and the whole register file (at two CTAs per SM they
// spill ~500 B of stores / ~1700 B of loads under the 128-register cap); 0: two CTAs per SM like the MLP kernels
#ifndef DGMK_TILE_DGM_ONE_CTA
#define DGMK_TILE_DGM_ONE_CTA 1
#endif
enum { PROB_HEAT = 0, PROB_ODE = 1, PROB_FRED = 2 };   // ODE covers simple_ode and FitzHugh-Nagumo (OdeArgs::fhn)

struct TileParams {
  NetDims n; PackedLayout pl;
  const float* Wp;        // packed weights (plain copy), global memory
  float* slots;           // [nslots][g_total] zero-initialised partial gradient accumulators
  int64_t B;              // points of this launch
  int32_t P;              // points per tile
  int32_t nslots_per_cta;
  int32_t flush_tiles;    // tiles per FP32 accumulation segment (FLUSH_TILES; a diagnostic knob lowers it in tests)
  int32_t w_smem, g_smem; // stage the weights / keep the accumulators in shared memory
  uint32_t w_floats, g_floats, lp_floats, coord_floats, ip_floats, tile_bytes;
  int32_t J;              // Fredholm: quadrature nodes per sub-tile (rows per sub-tile = P * J)
  HeatArgs heat;
  OdeArgs ode;
  FredArgs fred;
  long long* prof; int32_t prof_n;   // stage timeline of CTA 0 (diagnostic; nullptr = off)
};

}  // namespace tk
}  // namespace dgmk
