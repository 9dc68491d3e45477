// Layer-wise jet pipeline: forward pass (value + input-derivative channels through
// every layer, a-form stash in HBM), reverse pass (adjoint of the jet program),
// weight-gradient contractions, and the four fused training steps.
//
// gemm_nn(A, lda, B, ldb, Bt, ldbt, C, ...) receives the weight operand in BOTH packed
// orientations (B = [K,N] for the FFMA tile, Bt = [N,K] K-major for the tcgen05 tile).
//
// Templated on a backend that provides the heavy primitives (GEMM tiles, element-
// wise launches, reductions).  The product backend is CUDA (dgmk_cuda.cu); the
// test-only host harness (tests/host_emul) instantiates the same orchestration with
// plain loops so that buffer offsets, ordering and scaling can be validated against
// the oracle without a GPU.  There is no CPU path in the shipped library.
//
// Reference path replaced: heat.py:50-95 + :136-141, simple_ode.py:41-63 + :96-104,
// fitzhugh_nagumo.py:53-97 + :135-143, fredholm.py:47-74 + :104-109.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include "dgmk_layout.h"
#include "dgmk_math.h"
#include "dgmk_ops.h"

namespace dgmk {

// ---- workspace carving ----------------------------------------------------------
struct Carver {
  char* base; size_t off, cap; bool ok;
  DGMK_HD Carver(void* b, size_t c) : base((char*)b), off(0), cap(c), ok(true) {}
  DGMK_HD float* take(int64_t nfloats) {
    size_t bytes = ((size_t)nfloats * 4 + 255) / 256 * 256;
    if (off + bytes > cap) { ok = false; return nullptr; }
    float* p = (float*)(base + off);
    off += bytes;
    return p;
  }
};
DGMK_HD size_t carve_bytes(int64_t nfloats) { return ((size_t)nfloats * 4 + 255) / 256 * 256; }

// Buffers of one pass (one set of rows with one channel set)
struct PassBufs {
  XSrc xs; int64_t rows; int cs, C; int64_t M;
  float* E;            // [M][4]
  float* S[MAX_L + 1]; // layer states / MLP activations, [M][Hp]
  float* G[MAX_L];     // a-form gates [M][NG*Hp]
  float* SR[MAX_L];    // DGM: s*R [M][Hp]
  float* U;            // [M][4] output jets
  float* UB;           // [M][4] output cotangents
};
DGMK_HD int64_t pass_floats_per_row(const NetDims& n, int C) {
  // E + states + gates + SR + U + UB (each carved separately; rounding slack added by caller)
  int64_t Hp = n.Hp;
  int64_t f = 4 + (n.L + 1) * Hp + n.L * n.NG * Hp + (n.is_dgm() ? n.L * Hp : 0) + 4 + 4;
  return f * C;
}
// reverse scratch, shared by all passes of a step (sized for the largest M)
struct RevBufs { float* SBa; float* SBb; float* AB; float* SRB; };
DGMK_HD int64_t rev_floats_per_row(const NetDims& n, int C) {
  int64_t Hp = n.Hp;
  return (2 * Hp + n.NG * Hp + (n.is_dgm() ? Hp : 0)) * C;
}
constexpr int64_t PART_FLOATS_MIN = 1 << 22;   // 16 MB: 2048 partials of a hidden-size-32 layer
constexpr int TILE_MAX_HP = 128;         // hidden sizes the resident-tile step covers (above 64: small batches, and only when forced)
constexpr int TILE_WIDE_HP = 64;         // ... above this only batches of <= TILE_WIDE_ROWS rows (the reference's own regime)
constexpr int64_t TILE_WIDE_ROWS = 256;
constexpr int64_t TILE_SLOTS = 320;      // per-CTA partial gradient slots it may ask for (>= 2 per SM)
constexpr int64_t TILE_SLOTS_WIDE = 64;  // hidden sizes above TILE_WIDE_HP: one slot per 4-point tile of a 256-row batch
inline int64_t part_floats(const NetDims& n) {
  // gemm_tn partials: >= 256 splits x [3Hp, Hp]; wcolsum partials: <= 512 blocks x 4 x 4Hp
  int64_t a = 256LL * (3 * n.Hp * n.Hp + 4 * 3 * n.Hp), b = 512LL * 4 * 4 * n.Hp;
  int64_t m = a > b ? a : b;
  if (n.Hp <= TILE_MAX_HP) {   // resident-tile step: one partial copy of the packed gradient per slot
    PackedLayout pl; make_packed_layout(n, &pl);
    const int64_t slots = n.Hp > TILE_WIDE_HP ? TILE_SLOTS_WIDE : TILE_SLOTS;
    if (slots * pl.g_total > m) m = slots * pl.g_total;
  }
  return m > PART_FLOATS_MIN ? m : PART_FLOATS_MIN;
}

struct Ctx {
  NetDims n; PackedLayout pl; SegTable pack, grad;
  float* Wp;    // packed weights
  float* Gp;    // packed gradient accumulators (Gp[pl.g_acc] = loss)
  float* part;  // reduction partials
  int64_t part_n;
  float* Lp;    // per-point loss contributions
};

DGMK_HD bool carve_pass(Carver& cv, const NetDims& n, PassBufs* pb, int64_t rows, int cs) {
  pb->rows = rows; pb->cs = cs; pb->C = cs_channels(cs); pb->M = rows * pb->C;
  const int64_t M = pb->M, Hp = n.Hp;
  pb->E = cv.take(M * 4);
  for (int l = 0; l <= n.L; ++l) pb->S[l] = cv.take(M * Hp);
  for (int l = 0; l < n.L; ++l) {
    pb->G[l] = cv.take(M * n.NG * Hp);
    pb->SR[l] = n.is_dgm() ? cv.take(M * Hp) : nullptr;
  }
  pb->U = cv.take(M * 4);
  pb->UB = cv.take(M * 4);
  return cv.ok;
}
DGMK_HD size_t pass_bytes(const NetDims& n, int64_t rows, int cs) {
  int64_t C = cs_channels(cs), M = rows * C, Hp = n.Hp;
  size_t b = carve_bytes(M * 4) * 3 + carve_bytes(M * Hp) * (n.L + 1) + carve_bytes(M * n.NG * Hp) * n.L;
  if (n.is_dgm()) b += carve_bytes(M * Hp) * n.L;
  return b;
}
// inplace (resident-tile step: shared memory is the scarce resource): the reverse pass overwrites the forward
// stash as it consumes it -- pre-activation cotangents over the a-form gates they are computed from, (s*R)bar over
// s*R, one state-cotangent buffer instead of a ping-pong pair -- and needs a single [M, Hp] buffer of its own
DGMK_HD bool carve_rev(Carver& cv, const NetDims& n, RevBufs* rb, int64_t Mmax, bool inplace = false) {
  rb->SBa = cv.take(Mmax * n.Hp);
  if (inplace) { rb->SBb = rb->SBa; rb->AB = nullptr; rb->SRB = nullptr; return cv.ok; }
  rb->SBb = cv.take(Mmax * n.Hp);
  rb->AB = cv.take(Mmax * n.NG * n.Hp);
  rb->SRB = n.is_dgm() ? cv.take(Mmax * n.Hp) : nullptr;
  return cv.ok;
}
DGMK_HD size_t rev_bytes(const NetDims& n, int64_t Mmax, bool inplace = false) {
  if (inplace) return carve_bytes(Mmax * n.Hp);
  size_t b = carve_bytes(Mmax * n.Hp) * 2 + carve_bytes(Mmax * n.NG * n.Hp);
  if (n.is_dgm()) b += carve_bytes(Mmax * n.Hp);
  return b;
}
inline bool carve_ctx(Carver& cv, Ctx* c, int64_t max_points) {
  c->Wp = cv.take(c->pl.w_total * 3);  // plain | tf32-hi | tf32-lo
  c->Gp = cv.take(c->pl.g_total);
  c->part_n = part_floats(c->n);
  c->part = cv.take(c->part_n);
  c->Lp = cv.take(max_points);
  return cv.ok;
}
inline size_t ctx_bytes(const NetDims& n, const PackedLayout& pl, int64_t max_points) {
  return carve_bytes(pl.w_total * 3) + carve_bytes(pl.g_total) + carve_bytes(part_floats(n)) + carve_bytes(max_points);
}

// ---- dispatch helpers --------------------------------------------------------------
// A backend says which channel sets / activations / network families it is compiled for (BackendTraitsAll:
// everything -- the host-launched backends).  The persistent tile kernels (dgmk_tile.cuh) run this same
// orchestration INSIDE a kernel, instantiated for exactly the channel sets and activation of one problem, so that
// the run-time switches below compile to the one live branch.
struct BackendTraitsAll {
  static constexpr bool kHasTile = false;   // resident-tile step (dgmk_tile.cuh): CUDA backend only
  // Stage grouping (resident-tile step: a stage boundary is a CTA barrier that costs ~2000 cycles of a ~45000-cycle
  // pass, profiles/r02_tile_timeline.txt).  Between nosync(true) and nosync(false) the element-wise / gemm_nn calls of
  // a backend that groups do NOT end in a barrier: the orchestration brackets only stages that neither read what an
  // earlier stage of the bracket writes nor write what it reads or writes, and the first stage after the bracket ends
  // the group with its own barrier.  Host-launched backends (stream order) ignore it.
  static constexpr bool kGroupsStages = false;
  DGMK_HD void nosync(bool) {}
  DGMK_HD bool inplace_rev() const { return false; }   // reverse pass overwrites the stash (carve_rev)
  // hook: a backend may copy the `rows` coordinate rows a pass reads to faster memory and re-point xs at the copy
  DGMK_HD void stage_coords(XSrc&, int64_t) const {}
  // > 0: the Fredholm step of this backend walks blocks of `fredholm_block_points()` points whose nodes are taken
  // `fredholm_block_nodes()` at a time (dgmk_steps.h fredholm_block) instead of whole chunks -- the test harness
  // switches it on to exercise, on the host, the body the resident-tile kernel runs
  DGMK_HD int fredholm_block_nodes() const { return 0; }
  DGMK_HD int fredholm_block_points() const { return 0; }
  static constexpr bool cs_on(int) { return true; }
  static constexpr bool act_on(int) { return true; }
  static constexpr bool mlp_on() { return true; }
  static constexpr bool dgm_on() { return true; }
};
#define DGMK_CS_CASE(id, T, CS, ...) \
  case id: if constexpr (BK::cs_on(id)) { using CS = T; __VA_ARGS__; } break;
#define DGMK_CS_SWITCH(cs, CS, ...)                                  \
  switch (cs) {                                                      \
    DGMK_CS_CASE(CS_V, CsV, CS, __VA_ARGS__)                         \
    DGMK_CS_CASE(CS_D1O1, CsD1O1, CS, __VA_ARGS__)                   \
    DGMK_CS_CASE(CS_HEAT, CsHeat, CS, __VA_ARGS__)                   \
    DGMK_CS_CASE(CS_D2O1, CsD2O1, CS, __VA_ARGS__)                   \
    DGMK_CS_CASE(CS_D1O2, CsD1O2, CS, __VA_ARGS__)                   \
    default: if constexpr (BK::cs_on(CS_D2O2)) { using CS = CsD2O2; __VA_ARGS__; } break; \
  }
#define DGMK_ACT_CASE(id, ACT, ...) \
  case id: if constexpr (BK::act_on(id)) { constexpr int ACT = id; __VA_ARGS__; } break;
#define DGMK_ACT_SWITCH(act, ACT, ...)                                       \
  switch (act) {                                                             \
    DGMK_ACT_CASE(ACT_RELU, ACT, __VA_ARGS__)                                \
    DGMK_ACT_CASE(ACT_SIGMOID, ACT, __VA_ARGS__)                             \
    DGMK_ACT_CASE(ACT_TANH, ACT, __VA_ARGS__)                                \
    default: if constexpr (BK::act_on(ACT_LEAKY)) { constexpr int ACT = ACT_LEAKY; __VA_ARGS__; } break; \
  }
// DGM stacks only ever use tanh (dgm_net) or relu (neural_networks.DGM)
#define DGMK_GACT_SWITCH(act, ACT, ...)                                                                   \
  if ((act) == ACT_TANH) { if constexpr (BK::act_on(ACT_TANH)) { constexpr int ACT = ACT_TANH; __VA_ARGS__; } } \
  else { if constexpr (BK::act_on(ACT_RELU)) { constexpr int ACT = ACT_RELU; __VA_ARGS__; } }

// nvcc: the methods below are __host__ __device__ templates; instantiated for a host-launched backend they
// call that backend's __host__ methods, which is fine because those instantiations never run on the device
#if defined(__CUDACC__)
#define DGMK_NOCHECK _Pragma("nv_exec_check_disable")
#define DGMK_HD_TEMPLATE DGMK_NOCHECK __host__ __device__
#define DGMK_HD_PLAIN __host__ __device__
#else
#define DGMK_NOCHECK
#define DGMK_HD_TEMPLATE
#define DGMK_HD_PLAIN
#endif

template <class BK, class CX = Ctx>
struct Pipeline {
  BK& bk; CX& c;
  DGMK_HD_TEMPLATE Pipeline(BK& b, CX& ctx) : bk(b), c(ctx) {}

  DGMK_HD const F4* inb() const { return (const F4*)(c.Wp + c.pl.inb); }
  DGMK_HD const F4* ub(int l) const { return (const F4*)(c.Wp + c.pl.ub[l]); }

  void pack(const float* theta) {
    bk.zero(c.Wp, (size_t)c.pl.w_total * 4 * 3);
    PackFn f; f.t = c.pack; f.theta = theta; f.packed = c.Wp; f.hl = c.pl.w_total;
    bk.hl_stride = c.pl.w_total;
    bk.note_bytes(4.0 * num_params(c.n) * 7.0);
    bk.ew(f, num_params(c.n));
  }
  void zero_grads() { bk.zero(c.Gp, (size_t)c.pl.g_total * 4); }
  void unpack(float* grad_theta, float* loss_out) {
    UnpackGradFn f; f.t = c.grad; f.gp = c.Gp; f.grad = grad_theta;
    if (grad_theta) bk.ew(f, num_params(c.n));
    if (loss_out) bk.copy(loss_out, c.Gp + c.pl.g_acc, 4);
  }

  // ---------------------------------------------------------------- forward
  DGMK_HD_TEMPLATE void forward(PassBufs& pb) {
    const NetDims& n = c.n;
    const int Hp = n.Hp;
    const int64_t M = pb.M, R = pb.rows;
    DGMK_CS_SWITCH(pb.cs, CS, {
      const double unit = 4.0 * (double)M * Hp;   // one [M, Hp] FP32 matrix
      ExtInputFn<CS> fe; fe.xs = pb.xs; fe.E = pb.E;
      bk.note_bytes(R * (4.0 * n.d + 16.0 * pb.C));
      bk.nosync(true);    // E is first read by the reverse pass; the input layer reads the coordinates itself
      bk.ew(fe, R);
      bk.nosync(false);
      DGMK_ACT_SWITCH(n.in_act(), ACT, {
        InputFwdFn<CS, ACT> f; f.xs = pb.xs; f.inb = inb(); f.S0 = pb.S[0]; f.Hp = Hp;
        bk.note_bytes(unit);
        bk.ew4(f, R * Hp);
      })
      for (int l = 0; l < n.L; ++l) {
        if (!n.is_dgm()) { if constexpr (BK::mlp_on()) {
          if (bk.lane_ok(Hp, pb.cs)) {   // GEMM + bias + activation in one kernel
            DGMK_ACT_SWITCH(n.act, ACT, {
              bk.template mlp_fwd_fused<CS, ACT>(pb.S[l], pb.G[l], ub(l), pb.S[l + 1], c.Wp + c.pl.wb[l], Hp, M);
            })
            continue;
          }
          bk.gemm_nn(pb.S[l], Hp, c.Wp + c.pl.wf[l], Hp, c.Wp + c.pl.wb[l], Hp, pb.G[l], Hp, M, Hp, Hp, false);
          DGMK_ACT_SWITCH(n.act, ACT, {
            MlpActFn<CS, ACT> f; f.G = pb.G[l]; f.ub = ub(l); f.Yn = pb.S[l + 1]; f.Hp = Hp;
            bk.note_bytes(3 * unit);
            bk.ew(f, R * Hp);
          })
        } } else if constexpr (BK::dgm_on()) {
          if (bk.lane_ok(Hp, pb.cs)) {   // two kernels per layer: [Z|G|R] + s*R, then H + state update
            DGMK_GACT_SWITCH(n.gate_act(), ACT, {
              bk.template dgm_fwd_fused<CS, ACT>(pb.xs, pb.S[l], pb.G[l], ub(l), pb.SR[l], pb.S[l + 1], c.Wp + c.pl.wb[l], Hp, M);
            })
            continue;
          }
          bk.gemm_nn(pb.S[l], Hp, c.Wp + c.pl.wf[l], 3 * Hp, c.Wp + c.pl.wb[l], Hp, pb.G[l], 4 * Hp, M, 3 * Hp, Hp, false);
          DGMK_GACT_SWITCH(n.gate_act(), ACT, {
            DgmFwd1Fn<CS, ACT> f; f.xs = pb.xs; f.A4 = pb.G[l]; f.ub = ub(l); f.S = pb.S[l]; f.SR = pb.SR[l]; f.Hp = Hp;
            bk.note_bytes(8 * unit);
            bk.ew(f, R * Hp);
          })
          bk.gemm_nn(pb.SR[l], Hp, c.Wp + c.pl.wfh[l], Hp, c.Wp + c.pl.wb[l] + (int64_t)3 * Hp * Hp, Hp, pb.G[l] + 3 * Hp, 4 * Hp, M, Hp, Hp,
                     false);
          DGMK_GACT_SWITCH(n.gate_act(), ACT, {
            DgmFwd2Fn<CS, ACT> f; f.xs = pb.xs; f.A4 = pb.G[l]; f.ub = ub(l); f.S = pb.S[l]; f.Sn = pb.S[l + 1]; f.Hp = Hp;
            bk.note_bytes(6 * unit);
            bk.ew(f, R * Hp);
          })
        }
      }
    })
    bk.rowdot(pb.S[n.L], Hp, c.Wp + c.pl.outw, c.Wp + c.pl.outb, pb.U, M, Hp, n.o, pb.C);
  }

  // ---------------------------------------------------------------- reverse
  // pb.UB holds the output cotangents; gradients are ACCUMULATED into c.Gp.  loss_rows > 0: the pass's loss rows
  // (c.Lp[0 .. loss_rows)) are added to the loss accumulator first (add_loss) -- a backend that groups stages does
  // that sum, the two output-layer column sums and the output-layer adjoint in ONE stage.
  DGMK_HD_TEMPLATE void reverse(PassBufs& pb, RevBufs& rb, int64_t loss_rows = 0) {
    const NetDims& n = c.n;
    const int Hp = n.Hp;
    const int64_t M = pb.M, R = pb.rows;
    float* Gp = c.Gp;
    const double unit = 4.0 * (double)M * Hp;
    const bool ip = bk.inplace_rev();
    float* SBn = rb.SBa;  // cotangent of the current layer's output
    float* SBp = rb.SBb;  // (in-place mode: the same buffer -- every stage reads an element before it overwrites it)
    OutRevFn fo; fo.UB = pb.UB; fo.outw = c.Wp + c.pl.outw; fo.SB = SBn; fo.Hp = Hp; fo.o = n.o;
    // fused MLP reverse (hidden size 128, see the layer loop): the output-layer adjoint is part of the top layer's stage
    const bool fz = !ip && !n.is_dgm() && bk.lane_ok(Hp, pb.cs);
    if constexpr (BK::kGroupsStages) {
      bk.nosync(true);     // SBn is a buffer of its own: nothing the column sums read or write
      bk.ew4(fo, M * Hp);
      bk.nosync(false);
      bk.wcolsum3(loss_rows > 0 ? c.Lp : nullptr, loss_rows, Gp + c.pl.g_acc,      // loss sum
                  pb.S[n.L], Hp, pb.UB, M, Gp + c.pl.g_outw,                        // grad W_out = UB^T S_L
                  pb.E, Gp + c.pl.g_outb);                                          // grad b_out = UB^T E
    } else {
      if (loss_rows > 0) add_loss(loss_rows);
      // output layer: grad W_out = UB^T S_L, grad b_out = sum of value-row cotangents
      bk.wcolsum_acc(pb.S[n.L], Hp, Hp, pb.UB, M, Gp + c.pl.g_outw, c.part, c.part_n);
      bk.wcolsum_acc(pb.UB, 4, 4, pb.E, M, Gp + c.pl.g_outb, c.part, c.part_n);
      if (!fz) {
        bk.note_bytes(unit + 16.0 * M);
        bk.ew4(fo, M * Hp);
      }
    }
    float* ABin_fused = nullptr;
    DGMK_CS_SWITCH(pb.cs, CS, {
      float* fcur = nullptr; float* fnxt = nullptr;   // fused MLP reverse: abar_l ping-pong
      for (int l = n.L - 1; l >= 0; --l) {
        float* AB = ip ? pb.G[l] : rb.AB;      // pre-activation cotangents
        if (!n.is_dgm()) { if constexpr (BK::mlp_on()) {
          // Fused path (hidden size 128): the top layer's activation adjoint forms the output-layer cotangent in
          // registers (OutMlpRevFn); below it the data gradient and the activation adjoint of the layer underneath
          // are ONE launch (MlpRevEpi; at the bottom InputRevEpi for the input layer) -- abar_l ping-pongs between
          // rb.AB and rb.SBb and no cotangent y bar ever visits HBM.
          if (fz) {
            if (l == n.L - 1) { fcur = rb.AB; fnxt = rb.SBb; }
            AB = fcur;
          }
          if (fz && l == n.L - 1) {
            DGMK_ACT_SWITCH(n.act, ACT, {
              OutMlpRevFn<CS, ACT> f; f.UB = pb.UB; f.outw = c.Wp + c.pl.outw; f.G = pb.G[l]; f.AB = AB; f.Hp = Hp; f.o = n.o;
              bk.note_bytes(2 * unit + 16.0 * M);
              bk.ew(f, R * Hp);
            })
          } else if (!fz) {
            DGMK_ACT_SWITCH(n.act, ACT, {
              MlpRevFn<CS, ACT> f; f.G = pb.G[l]; f.YB = SBn; f.AB = AB; f.Hp = Hp;
              bk.note_bytes(3 * unit);
              bk.ew(f, R * Hp);
            })
          }
          if constexpr (BK::kGroupsStages) {   // the data gradient (AB -> SBp) and the weight gradient (AB, S[l] -> Gp): one stage
            bk.nosync(true);
            bk.gemm_nn(AB, Hp, c.Wp + c.pl.wb[l], Hp, c.Wp + c.pl.wf[l], Hp, SBp, Hp, M, Hp, Hp, false);
            bk.nosync(false);
            bk.gemm_tn_acc(AB, Hp, pb.S[l], Hp, Gp + c.pl.g_w[l], Hp, Hp, M, pb.E, Gp + c.pl.g_ub[l], Hp, c.part, c.part_n);
          } else {
          // grad W = Abar^T Y_prev, and grad b (row 2 of Abar^T E) in the same pass
          bk.gemm_tn_acc(AB, Hp, pb.S[l], Hp, Gp + c.pl.g_w[l], Hp, Hp, M, pb.E, Gp + c.pl.g_ub[l], Hp, c.part, c.part_n);
          if (fz) {
            if (l > 0) {
              DGMK_ACT_SWITCH(n.act, ACT, {
                bk.template mlp_rev_fused<CS, ACT>(AB, c.Wp + c.pl.wf[l], pb.G[l - 1], fnxt, Hp, M);
              })
              float* t2 = fcur; fcur = fnxt; fnxt = t2;
            } else {
              DGMK_ACT_SWITCH(n.in_act(), ACT, {
                bk.template input_rev_fused<CS, ACT>(AB, c.Wp + c.pl.wf[l], inb(), pb.S[0], fnxt, Hp, M);
              })
              ABin_fused = fnxt;
            }
            continue;   // (no state cotangents on this path)
          }
          if (bk.lane_ok(Hp, CS_V)) bk.lane_store(AB, Hp, c.Wp + c.pl.wf[l], SBp, Hp, Hp, M);
          else bk.gemm_nn(AB, Hp, c.Wp + c.pl.wb[l], Hp, c.Wp + c.pl.wf[l], Hp, SBp, Hp, M, Hp, Hp, false);
          }
        } } else if constexpr (BK::dgm_on()) {
          // fused path (hidden size 128): grad[U | b] is formed where the pre-activation cotangents are
          // produced (input_map_adj), so the weight-gradient passes carry no A^T E work
          const bool fe = bk.lane_ok(Hp, pb.cs);
          float* gub = Gp + c.pl.g_ub[l];
          float* SRB = ip ? pb.SR[l] : rb.SRB;   // (s*R)bar; in place once grad W_h has consumed s*R
          const float* Ew = fe ? nullptr : pb.E;
          DGMK_GACT_SWITCH(n.gate_act(), ACT, {
            DgmRev1Fn<CS, ACT> f; f.A4 = pb.G[l]; f.S = pb.S[l]; f.SBn = SBn; f.AB4 = AB; f.SBp = SBp; f.Hp = Hp;
            if (fe) bk.template dgm_rev1_e<CS>(f, pb.xs, R, gub, c.part, c.part_n);
            else { bk.note_bytes(9 * unit); bk.ew(f, R * Hp); }
          })
          // grad W_h = abar_H^T (s*R): abar_H is final here, and s*R is dead afterwards
          bk.gemm_tn_acc(AB + 3 * Hp, 4 * Hp, pb.SR[l], Hp, Gp + c.pl.g_w[l] + (int64_t)3 * Hp * Hp, Hp, Hp, M, Ew, gub + 3 * Hp,
                         4 * Hp, c.part, c.part_n);
          // (s*R)bar = abar_H W_h, then the R-gate adjoint (one kernel on the fused path)
          if (fe) {
            DGMK_GACT_SWITCH(n.gate_act(), ACT, {
              bk.template dgm_rev2_fused<CS, ACT>(pb.G[l], pb.S[l], AB, SBp, c.Wp + c.pl.wfh[l], Hp, M);
            })
            // grad[U_r | b_r] = abar_R^T E: a column-sum pass over abar_R (1 unit)
            bk.wcolsum_acc(AB + 2 * Hp, 4 * Hp, Hp, pb.E, M, gub + 2 * Hp, c.part, c.part_n, 4 * Hp);
          } else {
            bk.gemm_nn(AB + 3 * Hp, 4 * Hp, c.Wp + c.pl.wb[l] + (int64_t)3 * Hp * Hp, Hp, c.Wp + c.pl.wfh[l], Hp, SRB, Hp, M, Hp,
                       Hp, false);
            DGMK_GACT_SWITCH(n.gate_act(), ACT, {
              DgmRev2Fn<CS, ACT> f; f.A4 = pb.G[l]; f.S = pb.S[l]; f.SRB = SRB; f.AB4 = AB; f.SBp = SBp; f.Hp = Hp;
              bk.note_bytes(6 * unit);
              bk.ew(f, R * Hp);
            })
          }
          // s bar += [abar_Z | abar_G | abar_R] [W_z; W_g; W_r]  (grouping backends: same stage as the weight gradient
          // below -- it reads AB and S[l], this writes SBp)
          bk.nosync(true);
          bk.gemm_nn(AB, 4 * Hp, c.Wp + c.pl.wb[l], Hp, c.Wp + c.pl.wf[l], 3 * Hp, SBp, Hp, M, Hp, 3 * Hp, true);
          bk.nosync(false);
          // weight gradient of the Z, G, R gates; off the fused path grad[U | b] = Abar^T E rides along
          bk.gemm_tn_acc(AB, 4 * Hp, pb.S[l], Hp, Gp + c.pl.g_w[l], 3 * Hp, Hp, M, Ew, gub, 4 * Hp, c.part, c.part_n);
        }
        float* t = SBn; SBn = SBp; SBp = t;
      }
      if (ABin_fused) {
        bk.wcolsum_acc(ABin_fused, Hp, Hp, pb.E, M, Gp + c.pl.g_inb, c.part, c.part_n);
      } else {
      DGMK_ACT_SWITCH(n.in_act(), ACT, {
        float* ABin = ip ? SBn : rb.AB;
        InputRevFn<CS, ACT> f; f.inb = inb(); f.S0 = pb.S[0]; f.SB = SBn; f.AB = ABin; f.Hp = Hp; f.ldab = Hp;
        bk.note_bytes(2 * unit + 4.0 * (double)R * Hp);
        bk.ew4(f, R * Hp);
        bk.wcolsum_acc(ABin, Hp, Hp, pb.E, M, Gp + c.pl.g_inb, c.part, c.part_n);
      })
      }
    })
  }

  DGMK_HD_TEMPLATE void add_loss(int64_t rows) {
    bk.wcolsum_acc(c.Lp, 1, 1, nullptr, rows, c.Gp + c.pl.g_acc, c.part, c.part_n);
  }
};

DGMK_HD XSrc xsrc1(const float* p, int64_t rows, int d) {
  XSrc x; x.p[0] = p; x.p[1] = x.p[2] = nullptr; x.block_rows = rows > 0 ? rows : 1; x.block_stride = 0; x.nptr = 1; x.d = d;
  return x;
}

}  // namespace dgmk
