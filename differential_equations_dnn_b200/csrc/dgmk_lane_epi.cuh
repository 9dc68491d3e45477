// Fused epilogues of the units-on-lanes GEMM (dgmk_lane_gemm.cuh): the element-wise jet stages of
// dgmk_ops.h applied to the GEMM result while it is still in registers.  An epilogue thread owns
// unit j of gate `gate` for the CTA's whole life and walks the tile in groups of 8 consecutive
// GEMM rows = 8 / C collocation points with all their channels.
//
//   Const  init(gate, j)                          per-thread constants (input-map weights)
//   void   tile(t, k, row0, M, lane)              once per tile: coordinates of the thread's EPI_ROWS rows
//   void   prefetch(pre, t, k, row0, cg, M)       ISSUE every global load group cg needs
//   void   apply(pre, st, t, cg, k, row0, M, acc) acc[q] = W[j,:] . X[row0 + q,:]; math + stores
//   State / finish(st, k, slot)                   per-thread state across tiles, flushed at the end
//
// The split matters: only 8 epilogue warps (2 per scheduler) live on an SM, so latency has to be
// hidden inside each thread: the loads of a group are in flight while the previous group is being
// computed (and while the tile's MMAs still run), and apply() first computes ALL points of the
// group in straight-line code (independent tanh / jet chains that the compiler interleaves) and
// only then issues the (predicated) stores.  With load -> math -> store per point in sequence the
// epilogue was latency-bound at ~1300 cycles per point (measured, profiles/r01_notes.md).  The arithmetic per (point, unit) is the same sequence
// of operations as the stand-alone functors (DgmFwd1Fn, DgmFwd2Fn, MlpActFn, DgmRev2Fn), so both
// paths agree to the last bit given the same GEMM result.
#pragma once
#include "dgmk_ops.h"
#include "dgmk_lane_gemm.cuh"

namespace dgmk {
namespace lg {

constexpr int HP = 128;              // hidden size of every fused kernel (= K of the GEMM)
constexpr int64_t LD4 = 4 * HP;      // row pitch of the [M, 4*Hp] a-form / cotangent buffers

__device__ __forceinline__ float ldg_f(const float* p) { return __ldg(p); }

// Coordinates of the points of a warp's EPI_ROWS rows of the tile.  Lane l fetches point l once per tile
// (one XSrc::at -- an integer division -- per lane instead of one per point and lane); the groups
// then pick their points up by shuffle.
template <int C>
struct XTile {
  float x0, x1;
  __device__ __forceinline__ void load(const XSrc& xs, int64_t row0, int64_t M, int lane) {
    x0 = 0.f; x1 = 0.f;
    const int64_t r = row0 + (int64_t)lane * C;
    if (lane < EPI_ROWS / C && r < M) {
      const float* x = xs.at(r / C);
      x0 = ldg_f(x);
      if (xs.d > 1) x1 = ldg_f(x + 1);
    }
  }
  // point pp of group cg (8 rows per group)
  __device__ __forceinline__ void get(int cg, int pp, float& a, float& b) const {
    const int src = cg * (8 / C) + pp;
    a = __shfl_sync(0xffffffffu, x0, src);
    b = __shfl_sync(0xffffffffu, x1, src);
  }
};
template <int C>
struct XPre {
  float x0[8 / C], x1[8 / C];
  __device__ __forceinline__ void load(const XTile<C>& t, int cg) {
#pragma unroll
    for (int pp = 0; pp < 8 / C; ++pp) t.get(cg, pp, x0[pp], x1[pp]);
  }
};
struct NoTile {};
// a (+)= U x + b and the tangent seeds, as add_input_map (x[1] is only read when d > 1)
template <class CS>
__device__ __forceinline__ void add_input_map_xy(float* a, const F4& u, float x0, float x1, int d) {
  const float x[2] = {x0, x1};
  add_input_map<CS>(a, u, x, d);
}
// row index clamped into [0, M): loads of the padding rows of the last tile stay in bounds
// (FULL: the whole half tile lies inside [0, M) -- no clamps, no store predicates, and every address
// of the tile folds into one base register plus immediates)
template <bool FULL>
__device__ __forceinline__ int64_t clampr(int64_t r, int64_t M) { return (FULL || r < M) ? r : M - 1; }

// Z, G, R = act(W s + U x + b) -> a-form; SR = s * R        (dgm_net.py:63-65)
template <class CS, int ACT>
struct DgmFwd1Epi {
  XSrc xs; float* A4; const F4* ub; const float* S; float* SR;
  static constexpr int C = CS::C;
  struct Const { F4 u; int gate, j; };
  using Tile = XTile<C>;
  struct State {};
  __device__ __forceinline__ void finish(const State&, const Const&, int) const {}
  struct Pre { float s[8]; };   // (the coordinates are picked up by shuffle inside apply(): they sit in registers, and
                                // a prefetched copy per group in flight cost the value-row kernels 16 registers each)
  __device__ __forceinline__ Const init(int gate, int j) const { Const k; k.u = ub[gate * HP + j]; k.gate = gate; k.j = j; return k; }
  __device__ __forceinline__ void tile(Tile& t, const Const&, int64_t row0, int64_t M, int lane) const { t.load(xs, row0, M, lane); }
  template <bool FULL>
  __device__ __forceinline__ void prefetch(Pre& p, const Tile& t, const Const& k, int64_t row0, int cg, int64_t M) const {
    if (k.gate == 2) {
#pragma unroll
      for (int q = 0; q < 8; ++q) p.s[q] = ldg_f(S + clampr<FULL>(row0 + q, M) * HP + k.j);
    }
  }
  template <bool FULL>
  __device__ __forceinline__ void apply(const Pre& p, State& st, const Tile& t, int cg, const Const& k, int64_t row0, int64_t M, const float (&acc)[8]) const {
    float af[8], sr[8];   // a-form rows and (R gate) s*R rows of the whole group
    XPre<C> px;
    px.load(t, cg);
#pragma unroll
    for (int pp = 0; pp < 8 / C; ++pp) {
      float a[C], y[C];
#pragma unroll
      for (int c = 0; c < C; ++c) a[c] = acc[pp * C + c];
      add_input_map_xy<CS>(a, k.u, px.x0[pp], px.x1[pp], xs.d);
      act_fwd<CS, ACT>(a, y);
      af[pp * C] = y[0];
#pragma unroll
      for (int c = 1; c < C; ++c) af[pp * C + c] = a[c];
      if (k.gate == 2) {
        float s[C], t[C];
#pragma unroll
        for (int c = 0; c < C; ++c) s[c] = p.s[pp * C + c];
        prod_fwd<CS>(s, y, t);
#pragma unroll
        for (int c = 0; c < C; ++c) sr[pp * C + c] = t[c];
      }
    }
    float* ag = A4 + row0 * LD4 + k.gate * HP + k.j;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (FULL || row0 + q < M) ag[q * LD4] = af[q];
    if (k.gate == 2) {
      float* so = SR + row0 * HP + k.j;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (FULL || row0 + q < M) so[q * HP] = sr[q];
    }
  }
};

// H = act(W (s*R) + U x + b) -> a-form; s' = (1-G)*H + Z*s      (dgm_net.py:66-67)
template <class CS, int ACT>
struct DgmFwd2Epi {
  XSrc xs; float* A4; const F4* ub; const float* S; float* Sn;
  static constexpr int C = CS::C;
  struct Const { F4 u; int j; };
  using Tile = XTile<C>;
  struct State {};
  __device__ __forceinline__ void finish(const State&, const Const&, int) const {}
  struct Pre { float z[8], g[8], s[8]; };
  __device__ __forceinline__ Const init(int, int j) const { Const k; k.u = ub[3 * HP + j]; k.j = j; return k; }
  __device__ __forceinline__ void tile(Tile& t, const Const&, int64_t row0, int64_t M, int lane) const { t.load(xs, row0, M, lane); }
  template <bool FULL>
  __device__ __forceinline__ void prefetch(Pre& p, const Tile& t, const Const& k, int64_t row0, int cg, int64_t M) const {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int64_t r = clampr<FULL>(row0 + q, M);
      const float* row = A4 + r * LD4 + k.j;
      p.z[q] = ldg_f(row);
      p.g[q] = ldg_f(row + HP);
      p.s[q] = ldg_f(S + r * HP + k.j);
    }
  }
  template <bool FULL>
  __device__ __forceinline__ void apply(const Pre& p, State& st, const Tile& t, int cg, const Const& k, int64_t row0, int64_t M, const float (&acc)[8]) const {
    float af[8], sn[8];
    XPre<C> px;
    px.load(t, cg);
#pragma unroll
    for (int pp = 0; pp < 8 / C; ++pp) {
      float a[C], h[C], z[C], g[C], s[C], t1[C], t2[C];
#pragma unroll
      for (int c = 0; c < C; ++c) { a[c] = acc[pp * C + c]; t1[c] = p.z[pp * C + c]; t2[c] = p.g[pp * C + c]; s[c] = p.s[pp * C + c]; }
      add_input_map_xy<CS>(a, k.u, px.x0[pp], px.x1[pp], xs.d);
      act_fwd<CS, ACT>(a, h);
      af[pp * C] = h[0];
#pragma unroll
      for (int c = 1; c < C; ++c) af[pp * C + c] = a[c];
      aform_to_jet<CS, ACT>(t1, z);
      aform_to_jet<CS, ACT>(t2, g);
#pragma unroll
      for (int c = 0; c < C; ++c) g[c] = -g[c];
      g[0] += 1.0f;  // 1 - G
      prod_fwd<CS>(g, h, t1);
      prod_fwd<CS>(z, s, t2);
#pragma unroll
      for (int c = 0; c < C; ++c) sn[pp * C + c] = t1[c] + t2[c];
    }
    float* row = A4 + row0 * LD4 + 3 * HP + k.j;
    float* so = Sn + row0 * HP + k.j;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (FULL || row0 + q < M) { row[q * LD4] = af[q]; so[q * HP] = sn[q]; }
  }
};

// MLP hidden layer: y = act(W y_prev + b) -> a-form in G, output jet in Yn   (neural_networks.py:242-243)
template <class CS, int ACT>
struct MlpActEpi {
  float* G; const F4* ub; float* Yn;
  static constexpr int C = CS::C;
  struct Const { float b; int j; };
  using Tile = NoTile;
  struct State {};
  __device__ __forceinline__ void finish(const State&, const Const&, int) const {}
  struct Pre {};
  __device__ __forceinline__ Const init(int, int j) const { Const k; k.b = ub[j].z; k.j = j; return k; }
  __device__ __forceinline__ void tile(Tile&, const Const&, int64_t, int64_t, int) const {}
  template <bool FULL>
  __device__ __forceinline__ void prefetch(Pre&, const Tile&, const Const&, int64_t, int, int64_t) const {}
  template <bool FULL>
  __device__ __forceinline__ void apply(const Pre&, State&, const Tile&, int, const Const& k, int64_t row0, int64_t M, const float (&acc)[8]) const {
    float af[8], yo[8];
#pragma unroll
    for (int pp = 0; pp < 8 / C; ++pp) {
      float a[C], y[C];
#pragma unroll
      for (int c = 0; c < C; ++c) a[c] = acc[pp * C + c];
      a[0] += k.b;
      act_fwd<CS, ACT>(a, y);
      af[pp * C] = y[0];
#pragma unroll
      for (int c = 1; c < C; ++c) af[pp * C + c] = a[c];
#pragma unroll
      for (int c = 0; c < C; ++c) yo[pp * C + c] = y[c];
    }
    float* g = G + row0 * HP + k.j;
    float* y = Yn + row0 * HP + k.j;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (FULL || row0 + q < M) { g[q * HP] = af[q]; y[q * HP] = yo[q]; }
  }
};

// MLP reverse: y bar of layer l-1 = abar_l W_l arrives from the GEMM; abar_{l-1} = act_adj(y bar, a-form of layer l-1)
// (adjoint of neural_networks.py:242-243; stand-alone form: lane_store + MlpRevFn -- the cotangent never visits HBM)
template <class CS, int ACT>
struct MlpRevEpi {
  const float* G; float* AB;
  static constexpr int C = CS::C;
  struct Const { int j; };
  using Tile = NoTile;
  struct State {};
  __device__ __forceinline__ void finish(const State&, const Const&, int) const {}
  struct Pre { float af[8]; };
  __device__ __forceinline__ Const init(int, int j) const { Const k; k.j = j; return k; }
  __device__ __forceinline__ void tile(Tile&, const Const&, int64_t, int64_t, int) const {}
  template <bool FULL>
  __device__ __forceinline__ void prefetch(Pre& p, const Tile&, const Const& k, int64_t row0, int, int64_t M) const {
#pragma unroll
    for (int q = 0; q < 8; ++q) p.af[q] = ldg_f(G + clampr<FULL>(row0 + q, M) * HP + k.j);
  }
  template <bool FULL>
  __device__ __forceinline__ void apply(const Pre& p, State&, const Tile&, int, const Const& k, int64_t row0, int64_t M, const float (&acc)[8]) const {
    float abo[8];
#pragma unroll
    for (int pp = 0; pp < 8 / C; ++pp) {
      float af[C], yb[C], ab[C];
#pragma unroll
      for (int c = 0; c < C; ++c) { af[c] = p.af[pp * C + c]; yb[c] = acc[pp * C + c]; }
      act_adj<CS, ACT>(yb, af, ab);
#pragma unroll
      for (int c = 0; c < C; ++c) abo[pp * C + c] = ab[c];
    }
    float* o = AB + row0 * HP + k.j;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (FULL || row0 + q < M) o[q * HP] = abo[q];
  }
};

// MLP reverse, bottom layer: y bar of the INPUT layer = abar_0 W_0 arrives from the GEMM; abar_in = act_adj(y bar, input
// a-form) with the a-form rebuilt like InputRevFn does (value from S0, tangents = the input weights)
template <class CS, int ACT>
struct InputRevEpi {
  const F4* inb; const float* S0; float* AB;
  static constexpr int C = CS::C;
  struct Const { F4 w; int j; };
  using Tile = NoTile;
  struct State {};
  __device__ __forceinline__ void finish(const State&, const Const&, int) const {}
  struct Pre { float s0[8 / C]; };
  __device__ __forceinline__ Const init(int, int j) const { Const k; k.w = inb[j]; k.j = j; return k; }
  __device__ __forceinline__ void tile(Tile&, const Const&, int64_t, int64_t, int) const {}
  template <bool FULL>
  __device__ __forceinline__ void prefetch(Pre& p, const Tile&, const Const& k, int64_t row0, int, int64_t M) const {
#pragma unroll
    for (int pp = 0; pp < 8 / C; ++pp) p.s0[pp] = ldg_f(S0 + clampr<FULL>(row0 + pp * C, M) * HP + k.j);
  }
  template <bool FULL>
  __device__ __forceinline__ void apply(const Pre& p, State&, const Tile&, int, const Const& k, int64_t row0, int64_t M, const float (&acc)[8]) const {
    float abo[8];
#pragma unroll
    for (int pp = 0; pp < 8 / C; ++pp) {
      float af[C], yb[C], ab[C];
      af[0] = p.s0[pp];
#pragma unroll
      for (int kk = 0; kk < CS::ND; ++kk) af[1 + kk] = (kk == 0) ? k.w.x : k.w.y;
#pragma unroll
      for (int qq = 0; qq < CS::NP; ++qq) af[1 + CS::ND + qq] = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) yb[c] = acc[pp * C + c];
      act_adj<CS, ACT>(yb, af, ab);
#pragma unroll
      for (int c = 0; c < C; ++c) abo[pp * C + c] = ab[c];
    }
    float* o = AB + row0 * HP + k.j;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (FULL || row0 + q < M) o[q * HP] = abo[q];
  }
};

// (s*R)bar = abar_H W_h arrives from the GEMM; abar_R = act_adj((sR)bar * s), s bar += (sR)bar * R
// (adjoint of dgm_net.py:65-66; stand-alone form: DgmRev2Fn)
template <class CS, int ACT>
struct DgmRev2Epi {
  const float* A4; const float* S; float* AB4; float* SBp;
  static constexpr int C = CS::C;
  struct Const { int j; };
  using Tile = NoTile;
  struct State {};
  __device__ __forceinline__ void finish(const State&, const Const&, int) const {}
  struct Pre { float afr[8], s[8], sbar[8]; };
  __device__ __forceinline__ Const init(int, int j) const { Const k; k.j = j; return k; }
  __device__ __forceinline__ void tile(Tile&, const Const&, int64_t, int64_t, int) const {}
  template <bool FULL>
  __device__ __forceinline__ void prefetch(Pre& p, const Tile&, const Const& k, int64_t row0, int, int64_t M) const {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int64_t r = clampr<FULL>(row0 + q, M);
      p.afr[q] = ldg_f(A4 + r * LD4 + 2 * HP + k.j);
      p.s[q] = ldg_f(S + r * HP + k.j);
      p.sbar[q] = SBp[r * HP + k.j];   // read-modify-write by this thread only
    }
  }
  // (grad[U_r | b_r] = abar_R^T E is NOT formed here: the coordinates and running sums it needs cost
  // ~10 registers the prefetch ring lives on -- measured +4 ms per step; a column-sum pass over abar_R
  // does it for 1 unit of HBM traffic instead)
  template <bool FULL>
  __device__ __forceinline__ void apply(const Pre& p, State&, const Tile&, int, const Const& k, int64_t row0, int64_t M, const float (&acc)[8]) const {
    float abo[8], sbo[8];
#pragma unroll
    for (int pp = 0; pp < 8 / C; ++pp) {
      float afr[C], rj[C], s[C], srb[C], sbar[C], yb[C], ab[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        afr[c] = p.afr[pp * C + c]; s[c] = p.s[pp * C + c];
        srb[c] = acc[pp * C + c]; sbar[c] = p.sbar[pp * C + c];
      }
      aform_to_jet<CS, ACT>(afr, rj);
      prod_adj<CS, false>(srb, s, yb);  // Rbar
      act_adj<CS, ACT>(yb, afr, ab);
      prod_adj<CS, true>(srb, rj, sbar);
#pragma unroll
      for (int c = 0; c < C; ++c) { abo[pp * C + c] = ab[c]; sbo[pp * C + c] = sbar[c]; }
    }
    float* orow = AB4 + row0 * LD4 + 2 * HP + k.j;
    float* so = SBp + row0 * HP + k.j;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (FULL || row0 + q < M) { orow[q * LD4] = abo[q]; so[q * HP] = sbo[q]; }
  }
};

}  // namespace lg
}  // namespace dgmk
