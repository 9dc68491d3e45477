// Parameter layouts: the reference's flat named_parameters() order ("theta") and
// the padded, GEMM-friendly packed layout the kernels read ("packed").
//
// theta order (checked against the executed reference in oracle/make_golden.py):
//   MLP                neural_networks.py:184-228  fc_in.{weight[H,d],bias[H]},
//                      layers.i.{weight[H,H],bias[H]}, fc_out.{weight[o,H],bias[o]}
//   dgm_net.DGM        dgm_net.py:75-101, layer :38-48   S_in, per layer
//                      (Z_wg, Z_ug, G_wz, G_uz, R_wr, R_ur, H_wh, H_uh), S_out
//   neural_networks.DGM neural_networks.py:134-160, layer :67-96  x_in, dgm1 (dead),
//                      per layer (Uz Ug Ur Uh [d,H], Wz Wg Wr Wh [H,H] used as s@W,
//                      bz bg br bh [1,H]), x_out
#pragma once
#include <stdint.h>
#include "dgmk_math.h"

namespace dgmk {

constexpr int MAX_L = 16;
constexpr int MAX_SEGS = 104;

struct NetDims {
  int kind, d, o, H, L, act;
  int Hp;  // H rounded up to a multiple of 32
  int NG;  // gates per layer: 1 (MLP) or 4 (DGM: Z, G, R, H)
  DGMK_HD bool is_dgm() const { return kind != KIND_MLP; }
  // activation used by the hidden stack (dgm_net: tanh; neural_networks.DGM: relu)
  DGMK_HD int gate_act() const { return kind == KIND_MLP ? act : (kind == KIND_DGM_LINEAR ? ACT_TANH : ACT_RELU); }
  // input layer: the stack's activation, except neural_networks.DGM(func="tanh"), whose x_in is followed by
  // tanh while its layers stay ReLU (neural_networks.py:145-147,172)
  DGMK_HD int in_act() const { return kind == KIND_DGM_RAW ? act : gate_act(); }
};

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

inline bool make_dims(int kind, int d, int o, int H, int L, int act, NetDims* nd, const char** err) {
  if (kind < 0 || kind > 2) { *err = "unknown net kind"; return false; }
  if (d < 1 || d > 2) { *err = "input_dim must be 1 or 2"; return false; }
  if (o < 1 || o > 4) { *err = "output_dim must be 1..4"; return false; }
  if (H < 1 || H > 512) { *err = "hidden_size must be 1..512"; return false; }
  if (L < 0 || L > MAX_L) { *err = "num_layers out of range"; return false; }
  if (kind != KIND_MLP && L > 8) { *err = "DGM num_layers must be <= 8"; return false; }
  if (act < 0 || act > 3) { *err = "unknown activation"; return false; }
  nd->kind = kind; nd->d = d; nd->o = o; nd->H = H; nd->L = L;
  if (kind == KIND_DGM_RAW && act != ACT_RELU && act != ACT_TANH) { *err = "neural_networks.DGM: func is relu or tanh"; return false; }
  nd->act = (kind == KIND_DGM_LINEAR) ? ACT_TANH : act;   // DGM_RAW: activation of the INPUT layer only
  nd->Hp = round_up(H, 32);
  nd->NG = (kind == KIND_MLP) ? 1 : 4;
  return true;
}

// one theta tensor [rows, cols] row-major and where element (r, c) lives
struct Seg {
  int32_t theta_off, n, cols;
  int32_t a_off, a_rs, a_cs;  // first packed destination (or -1)
  int32_t b_off, b_rs, b_cs;  // second packed destination (or -1)
};
struct SegTable {
  int32_t n;
  Seg s[MAX_SEGS];
};

struct PackedLayout {
  // packed weights (floats)
  int64_t inb, outw, outb;
  int64_t wf[MAX_L];   // MLP: [Hp,Hp] ; DGM: ZGR [Hp,3Hp]
  int64_t wfh[MAX_L];  // DGM: H gate [Hp,Hp]
  int64_t wb[MAX_L];   // [NG*Hp, Hp] rows = (gate, out unit), cols = in unit
  int64_t ub[MAX_L];   // [NG*Hp][4] = (U_j0, U_j1, b_j, 0)
  int64_t w_total;
  // packed gradient accumulators (floats); g_acc[0] = loss
  int64_t g_acc, g_inb, g_outw, g_outb;
  int64_t g_w[MAX_L];   // [NG*Hp, Hp]
  int64_t g_ub[MAX_L];  // [4][NG*Hp]
  int64_t g_total;
};

inline int64_t num_params(const NetDims& n) {
  int64_t H = n.H, d = n.d, o = n.o, L = n.L;
  if (n.kind == KIND_MLP) return H * d + H + L * (H * H + H) + o * H + o;
  if (n.kind == KIND_DGM_LINEAR) return H * d + H + L * 4 * (H * H + H + H * d) + o * H + o;
  return H * d + H + (L + 1) * 4 * (d * H + H * H + H) + o * H + o;
}

inline void make_packed_layout(const NetDims& n, PackedLayout* p) {
  const int64_t Hp = n.Hp, NG = n.NG;
  int64_t off = 0;
  auto take = [&](int64_t sz) { int64_t o = off; off += (sz + 31) / 32 * 32; return o; };
  p->inb = take(Hp * 4);
  p->outw = take(4 * Hp);
  p->outb = take(4);
  for (int l = 0; l < n.L; ++l) {
    if (n.is_dgm()) { p->wf[l] = take(Hp * 3 * Hp); p->wfh[l] = take(Hp * Hp); }
    else { p->wf[l] = take(Hp * Hp); p->wfh[l] = -1; }
    p->wb[l] = take(NG * Hp * Hp);
    p->ub[l] = take(NG * Hp * 4);
  }
  p->w_total = off;
  off = 0;
  p->g_acc = take(32);
  p->g_inb = take(4 * Hp);
  p->g_outw = take(4 * Hp);
  p->g_outb = take(16);
  for (int l = 0; l < n.L; ++l) { p->g_w[l] = take(NG * Hp * Hp); p->g_ub[l] = take(4 * NG * Hp); }
  p->g_total = off;
}

// Build both tables: `pack` maps theta -> packed weights (a = forward operand /
// small params, b = data-gradient operand); `grad` maps packed gradient -> theta.
inline bool make_seg_tables(const NetDims& n, const PackedLayout& p, SegTable* pack, SegTable* grad,
                            const char** err) {
  const int H = n.H, d = n.d, o = n.o, Hp = n.Hp, NG = n.NG;
  int32_t off = 0;
  pack->n = grad->n = 0;
  bool ok = true;
  auto add = [&](int rows, int cols, Seg ps, Seg gs) {
    if (pack->n >= MAX_SEGS) { ok = false; return; }
    ps.theta_off = gs.theta_off = off; ps.n = gs.n = rows * cols; ps.cols = gs.cols = cols;
    pack->s[pack->n++] = ps; grad->s[grad->n++] = gs;
    off += rows * cols;
  };
  auto S = [](int64_t a_off, int a_rs, int a_cs, int64_t b_off = -1, int b_rs = 0, int b_cs = 0) {
    Seg s{}; s.a_off = (int32_t)a_off; s.a_rs = a_rs; s.a_cs = a_cs;
    s.b_off = (int32_t)b_off; s.b_rs = b_rs; s.b_cs = b_cs; return s;
  };
  // input layer  W[H,d], b[H]
  add(H, d, S(p.inb, 4, 1), S(p.g_inb, 1, Hp));
  add(H, 1, S(p.inb + 2, 4, 0), S(p.g_inb + 2 * Hp, 1, 0));
  if (n.kind == KIND_MLP) {
    for (int l = 0; l < n.L; ++l) {
      add(H, H, S(p.wf[l], 1, Hp, p.wb[l], Hp, 1), S(p.g_w[l], Hp, 1));
      add(H, 1, S(p.ub[l] + 2, 4, 0), S(p.g_ub[l] + 2 * Hp, 1, 0));
    }
  } else if (n.kind == KIND_DGM_LINEAR) {
    for (int l = 0; l < n.L; ++l)
      for (int g = 0; g < 4; ++g) {
        int64_t wf = (g < 3) ? p.wf[l] + g * Hp : p.wfh[l];
        int ldf = (g < 3) ? 3 * Hp : Hp;
        add(H, H, S(wf, 1, ldf, p.wb[l] + (int64_t)g * Hp * Hp, Hp, 1), S(p.g_w[l] + (int64_t)g * Hp * Hp, Hp, 1));
        add(H, 1, S(p.ub[l] + (int64_t)g * Hp * 4 + 2, 4, 0), S(p.g_ub[l] + 2 * NG * Hp + g * Hp, 1, 0));
        add(H, d, S(p.ub[l] + (int64_t)g * Hp * 4, 4, 1), S(p.g_ub[l] + g * Hp, 1, NG * Hp));
      }
  } else {
    {  // dgm1: registered, never used (neural_networks.py:145): no pack, zero grad
      Seg dead = S(-1, 0, 0);
      add(1, 4 * (d * H + H * H + H), dead, dead);
    }
    for (int l = 0; l < n.L; ++l) {
      for (int g = 0; g < 4; ++g)  // U [d,H]
        add(d, H, S(p.ub[l] + (int64_t)g * Hp * 4, 1, 4), S(p.g_ub[l] + g * Hp, NG * Hp, 1));
      for (int g = 0; g < 4; ++g) {  // W [in,out]
        int64_t wf = (g < 3) ? p.wf[l] + g * Hp : p.wfh[l];
        int ldf = (g < 3) ? 3 * Hp : Hp;
        add(H, H, S(wf, ldf, 1, p.wb[l] + (int64_t)g * Hp * Hp, 1, Hp), S(p.g_w[l] + (int64_t)g * Hp * Hp, 1, Hp));
      }
      for (int g = 0; g < 4; ++g)  // b [1,H]
        add(1, H, S(p.ub[l] + (int64_t)g * Hp * 4 + 2, 0, 4), S(p.g_ub[l] + 2 * NG * Hp + g * Hp, 0, 1));
    }
  }
  add(o, H, S(p.outw, Hp, 1), S(p.g_outw, Hp, 1));
  add(o, 1, S(p.outb, 1, 0), S(p.g_outb + 2 * 4, 1, 0));
  if (!ok) { *err = "too many parameter tensors for the segment table"; return false; }
  if (off != num_params(n)) { *err = "internal: layout size mismatch"; return false; }
  return true;
}

}  // namespace dgmk
