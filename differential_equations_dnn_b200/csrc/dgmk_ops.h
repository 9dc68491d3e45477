// Element-wise stages of the jet pipeline, written as POD functors over a flat
// index so that the CUDA backend can launch them as kernels and the test-only
// host harness can loop over them.  One functor call handles one
// (collocation point, hidden unit) pair and owns ALL channels of that pair.
//
// Activations live in HBM as [rows*C, ld] row-major FP32, GEMM row = point*C + c.
// Gate outputs are stored in "a-form": channel 0 holds the OUTPUT value y,
// channels >=1 hold the PRE-activation tangents (what the adjoint needs:
// sigma'' and sigma''' terms multiply a_x, a_xx).
#pragma once
#include "dgmk_layout.h"
#include "dgmk_math.h"

namespace dgmk {

struct alignas(16) F4 { float x, y, z, w; };   // every F4 array in the workspace is 16-byte aligned

// Where the coordinates of row r of a pass come from: up to 3 separate arrays
// (heat companions X0 | XBD1 | XBD2) or one array walked in blocks with a stride
// (Fredholm nodes T[k, B]: block j = j-th Monte-Carlo draw).
struct XSrc {
  const float* p[3];
  int64_t block_rows;    // rows per block
  int64_t block_stride;  // floats between consecutive blocks when nptr == 1
  int32_t nptr, d;
  DGMK_HD const float* at(int64_t r) const {
    // separate arrays: at most three blocks, two comparisons instead of a division
    int64_t b = (nptr > 1) ? (int64_t)(r >= block_rows) + (int64_t)(r >= 2 * block_rows)
                           : ((r < block_rows) ? 0 : ((block_rows >> 31) == 0 ? idiv(r, (int32_t)block_rows) : r / block_rows));
    int64_t w = r - b * block_rows;
    // ternaries, not p[b]: a runtime index would spill the parameter array to local memory
    const float* base = (nptr > 1) ? (b == 0 ? p[0] : (b == 1 ? p[1] : p[2])) : p[0] + b * block_stride;
    return base + w * d;
  }
};

// ---- parameter packing --------------------------------------------------------
// packed = [plain | tf32-hi | tf32-lo], three copies of the same layout hl apart
struct PackFn {
  SegTable t; const float* theta; float* packed; int64_t hl;
  DGMK_HD void operator()(int64_t i) const {
    for (int s = 0; s < t.n; ++s) {
      const Seg& g = t.s[s];
      int32_t k = (int32_t)i - g.theta_off;
      if (k >= 0 && k < g.n) {
        int r = k / g.cols, c = k - r * g.cols;
        const float v = theta[i], hi = tf32_round(v), lo = v - hi;
        if (g.a_off >= 0) {
          const int64_t o = g.a_off + (int64_t)r * g.a_rs + (int64_t)c * g.a_cs;
          packed[o] = v; packed[o + hl] = hi; packed[o + 2 * hl] = lo;
        }
        if (g.b_off >= 0) {
          const int64_t o = g.b_off + (int64_t)r * g.b_rs + (int64_t)c * g.b_cs;
          packed[o] = v; packed[o + hl] = hi; packed[o + 2 * hl] = lo;
        }
        return;
      }
    }
  }
};
// grad_theta[i] = packed_grad[...] (dead parameters: 0)
struct UnpackGradFn {
  SegTable t; const float* gp; float* grad;
  DGMK_HD void operator()(int64_t i) const {
    for (int s = 0; s < t.n; ++s) {
      const Seg& g = t.s[s];
      int32_t k = (int32_t)i - g.theta_off;
      if (k >= 0 && k < g.n) {
        int r = k / g.cols, c = k - r * g.cols;
        grad[i] = (g.a_off >= 0) ? gp[g.a_off + (int64_t)r * g.a_rs + (int64_t)c * g.a_cs] : 0.f;
        return;
      }
    }
  }
};

// ---- E: extended inputs, [rows*C][4] = (x0|e0, x1|e1, 1|0, 0) -----------------
// grad[U | b] = Abar^T E  (SURVEY 7.1 "input map").
template <class CS>
struct ExtInputFn {
  XSrc xs; float* E;
  DGMK_HD void operator()(int64_t p) const {
    const float* x = xs.at(p);
    float* e = E + p * CS::C * 4;
    DGMK_SMEM(E);
    e[0] = x[0]; e[1] = (xs.d > 1) ? x[1] : 0.f; e[2] = 1.f; e[3] = 0.f;
#pragma unroll
    for (int c = 1; c < CS::C; ++c) {
      float* ec = e + c * 4;
      ec[0] = (c == 1) ? 1.f : 0.f;
      ec[1] = (c == 2 && CS::ND > 1) ? 1.f : 0.f;
      ec[2] = 0.f; ec[3] = 0.f;
    }
  }
};

// ---- input layer: s0 = act(W_in x + b) (dgm_net.py:112, neural_networks.py:241,172)
template <class CS, int ACT>
struct InputFwdFn {
  XSrc xs; const F4* inb; float* S0; int Hp;
  DGMK_HD void operator()(int64_t i) const {
    DGMK_SMEM(S0);
    int64_t p = idiv(i, Hp); int j = (int)(i - p * Hp);
    const float* x = xs.at(p);
    F4 w = inb[j];
    float a[CS::C], y[CS::C];
    a[0] = w.x * x[0] + w.z;
    if (xs.d > 1) a[0] += w.y * x[1];
#pragma unroll
    for (int k = 0; k < CS::ND; ++k) a[1 + k] = (k == 0) ? w.x : w.y;
#pragma unroll
    for (int q = 0; q < CS::NP; ++q) a[1 + CS::ND + q] = 0.f;
    act_fwd<CS, ACT>(a, y);
    float* o = S0 + (p * CS::C) * Hp + j;
#pragma unroll
    for (int c = 0; c < CS::C; ++c) o[(int64_t)c * Hp] = y[c];
  }
  // four consecutive units of one point (same arithmetic as operator(); 16-byte stores, the point's
  // coordinates and the index division once per four elements); k = p * (Hp/4) + j/4
  DGMK_HD void vec4(int64_t k) const {
    DGMK_SMEM(S0);
    const int q = Hp >> 2;
    int64_t p = idiv(k, q); int j = (int)(k - p * q) * 4;
    const float* x = xs.at(p);
    const float x0 = x[0], x1 = (xs.d > 1) ? x[1] : 0.f;
    float y[4][CS::C];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      F4 w = inb[j + u];
      float a[CS::C];
      a[0] = w.x * x0 + w.z;
      if (xs.d > 1) a[0] += w.y * x1;
#pragma unroll
      for (int kk = 0; kk < CS::ND; ++kk) a[1 + kk] = (kk == 0) ? w.x : w.y;
#pragma unroll
      for (int qq = 0; qq < CS::NP; ++qq) a[1 + CS::ND + qq] = 0.f;
      act_fwd<CS, ACT>(a, y[u]);
    }
    float* o = S0 + (p * CS::C) * Hp + j;
#pragma unroll
    for (int c = 0; c < CS::C; ++c) {
      F4 v; v.x = y[0][c]; v.y = y[1][c]; v.z = y[2][c]; v.w = y[3][c];
      *reinterpret_cast<F4*>(o + (int64_t)c * Hp) = v;
    }
  }
};
template <class CS, int ACT>
struct InputRevFn {
  const F4* inb; const float* S0; const float* SB; float* AB; int Hp; int64_t ldab;
  DGMK_HD void operator()(int64_t i) const {
    DGMK_SMEM(S0); DGMK_SMEM(SB); DGMK_SMEM(AB);
    int64_t p = idiv(i, Hp); int j = (int)(i - p * Hp);
    F4 w = inb[j];
    float af[CS::C], yb[CS::C], ab[CS::C];
    af[0] = S0[(p * CS::C) * Hp + j];
#pragma unroll
    for (int k = 0; k < CS::ND; ++k) af[1 + k] = (k == 0) ? w.x : w.y;
#pragma unroll
    for (int q = 0; q < CS::NP; ++q) af[1 + CS::ND + q] = 0.f;
#pragma unroll
    for (int c = 0; c < CS::C; ++c) yb[c] = SB[(p * CS::C + c) * Hp + j];
    act_adj<CS, ACT>(yb, af, ab);
#pragma unroll
    for (int c = 0; c < CS::C; ++c) AB[(p * CS::C + c) * ldab + j] = ab[c];
  }
  DGMK_HD void vec4(int64_t k) const {   // see InputFwdFn::vec4
    DGMK_SMEM(S0); DGMK_SMEM(SB); DGMK_SMEM(AB);
    const int q = Hp >> 2;
    int64_t p = idiv(k, q); int j = (int)(k - p * q) * 4;
    const F4 s0 = *reinterpret_cast<const F4*>(S0 + (p * CS::C) * Hp + j);
    F4 sb[CS::C];
#pragma unroll
    for (int c = 0; c < CS::C; ++c) sb[c] = *reinterpret_cast<const F4*>(SB + (p * CS::C + c) * Hp + j);
    float ab[4][CS::C];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      F4 w = inb[j + u];
      float af[CS::C], yb[CS::C];
      af[0] = u == 0 ? s0.x : (u == 1 ? s0.y : (u == 2 ? s0.z : s0.w));
#pragma unroll
      for (int kk = 0; kk < CS::ND; ++kk) af[1 + kk] = (kk == 0) ? w.x : w.y;
#pragma unroll
      for (int qq = 0; qq < CS::NP; ++qq) af[1 + CS::ND + qq] = 0.f;
#pragma unroll
      for (int c = 0; c < CS::C; ++c) yb[c] = u == 0 ? sb[c].x : (u == 1 ? sb[c].y : (u == 2 ? sb[c].z : sb[c].w));
      act_adj<CS, ACT>(yb, af, ab[u]);
    }
#pragma unroll
    for (int c = 0; c < CS::C; ++c) {
      F4 v; v.x = ab[0][c]; v.y = ab[1][c]; v.z = ab[2][c]; v.w = ab[3][c];
      *reinterpret_cast<F4*>(AB + (p * CS::C + c) * ldab + j) = v;
    }
  }
};

// ---- MLP hidden layer: y = act(W y_prev + b) (neural_networks.py:242-243) ------
template <class CS, int ACT>
struct MlpActFn {
  float* G; const F4* ub; float* Yn; int Hp;  // G: GEMM output -> a-form in place
  DGMK_HD void operator()(int64_t i) const {
    DGMK_SMEM(G); DGMK_SMEM(Yn);
    int64_t p = idiv(i, Hp); int j = (int)(i - p * Hp);
    float a[CS::C], y[CS::C];
    float* g = G + (p * CS::C) * Hp + j;
#pragma unroll
    for (int c = 0; c < CS::C; ++c) a[c] = g[(int64_t)c * Hp];
    a[0] += ub[j].z;
    act_fwd<CS, ACT>(a, y);
    g[0] = y[0];
    float* o = Yn + (p * CS::C) * Hp + j;
#pragma unroll
    for (int c = 0; c < CS::C; ++c) o[(int64_t)c * Hp] = y[c];
  }
};
template <class CS, int ACT>
struct MlpRevFn {
  const float* G; const float* YB; float* AB; int Hp;
  DGMK_HD void operator()(int64_t i) const {
    DGMK_SMEM(G); DGMK_SMEM(YB); DGMK_SMEM(AB);
    int64_t p = idiv(i, Hp); int j = (int)(i - p * Hp);
    float af[CS::C], yb[CS::C], ab[CS::C];
    int64_t base = (p * CS::C) * Hp + j;
#pragma unroll
    for (int c = 0; c < CS::C; ++c) { af[c] = G[base + (int64_t)c * Hp]; yb[c] = YB[base + (int64_t)c * Hp]; }
    act_adj<CS, ACT>(yb, af, ab);
#pragma unroll
    for (int c = 0; c < CS::C; ++c) AB[base + (int64_t)c * Hp] = ab[c];
  }
};

// ---- DGM layer (dgm_net.py:63-67; neural_networks.py:115-126) -------------------
// gate slots in the [rows*C, 4Hp] a-form buffer: 0=Z 1=G 2=R 3=H
template <class CS>
DGMK_HD void add_input_map(float* a, const F4& u, const float* x, int d) {
  a[0] += u.x * x[0] + u.z;
  if (d > 1) a[0] += u.y * x[1];
#pragma unroll
  for (int k = 0; k < CS::ND; ++k) a[1 + k] += (k == 0) ? u.x : u.y;
}
// adjoint of add_input_map for one point: g[0..1] += grad U[:, 0..1], g[2] += grad b.  Same sums as
// abar^T E over the point's rows (E: ExtInputFn), with the structural zeros and ones of E folded in
template <class CS>
DGMK_HD void input_map_adj(const float* ab, float x0, float x1, float* g) {
  g[0] = fmaf(ab[0], x0, g[0]);
  g[1] = fmaf(ab[0], x1, g[1]);
  g[2] += ab[0];
  if (CS::ND > 0) g[0] += ab[1];
  if (CS::ND > 1) g[1] += ab[2];
}
// stage 1: Z, G, R = act(W s + U x + b) in place (a-form), SR = s * R
template <class CS, int ACT>
struct DgmFwd1Fn {
  XSrc xs; float* A4; const F4* ub; const float* S; float* SR; int Hp;
  DGMK_HD void operator()(int64_t i) const {
    DGMK_SMEM(A4); DGMK_SMEM(S); DGMK_SMEM(SR);
    int64_t p = idiv(i, Hp); int j = (int)(i - p * Hp);
    const float* x = xs.at(p);
    const int64_t ld = 4 * (int64_t)Hp;
    float a[CS::C], y[CS::C];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      float* ag = A4 + (p * CS::C) * ld + g * Hp + j;
#pragma unroll
      for (int c = 0; c < CS::C; ++c) a[c] = ag[c * ld];
      add_input_map<CS>(a, ub[g * Hp + j], x, xs.d);
      act_fwd<CS, ACT>(a, y);
      ag[0] = y[0];
#pragma unroll
      for (int k = 0; k < CS::ND; ++k) ag[(1 + k) * ld] = a[1 + k];
    }
    // y now holds the R jet
    float s[CS::C], r[CS::C];
    int64_t sb = (p * CS::C) * Hp + j;
#pragma unroll
    for (int c = 0; c < CS::C; ++c) s[c] = S[sb + (int64_t)c * Hp];
    prod_fwd<CS>(s, y, r);
#pragma unroll
    for (int c = 0; c < CS::C; ++c) SR[sb + (int64_t)c * Hp] = r[c];
  }
};
// stage 2: H = act(W (s*R) + U x + b) in place; s' = (1-G)*H + Z*s
template <class CS, int ACT>
struct DgmFwd2Fn {
  XSrc xs; float* A4; const F4* ub; const float* S; float* Sn; int Hp;
  DGMK_HD void operator()(int64_t i) const {
    DGMK_SMEM(A4); DGMK_SMEM(S); DGMK_SMEM(Sn);
    int64_t p = idiv(i, Hp); int j = (int)(i - p * Hp);
    const float* x = xs.at(p);
    const int64_t ld = 4 * (int64_t)Hp;
    float a[CS::C], h[CS::C], z[CS::C], g[CS::C], s[CS::C], t1[CS::C], t2[CS::C];
    float* row = A4 + (p * CS::C) * ld + j;
#pragma unroll
    for (int c = 0; c < CS::C; ++c) a[c] = row[c * ld + 3 * Hp];
    add_input_map<CS>(a, ub[3 * Hp + j], x, xs.d);
    act_fwd<CS, ACT>(a, h);
    row[3 * Hp] = h[0];
#pragma unroll
    for (int k = 0; k < CS::ND; ++k) row[(1 + k) * ld + 3 * Hp] = a[1 + k];
#pragma unroll
    for (int c = 0; c < CS::C; ++c) { t1[c] = row[c * ld]; t2[c] = row[c * ld + Hp]; }
    aform_to_jet<CS, ACT>(t1, z);
    aform_to_jet<CS, ACT>(t2, g);
#pragma unroll
    for (int c = 0; c < CS::C; ++c) g[c] = -g[c];
    g[0] += 1.0f;  // 1 - G
    int64_t sb = (p * CS::C) * Hp + j;
#pragma unroll
    for (int c = 0; c < CS::C; ++c) s[c] = S[sb + (int64_t)c * Hp];
    prod_fwd<CS>(g, h, t1);
    prod_fwd<CS>(z, s, t2);
#pragma unroll
    for (int c = 0; c < CS::C; ++c) Sn[sb + (int64_t)c * Hp] = t1[c] + t2[c];
  }
};
// reverse stage 1: from s'bar -> abar_H (slot 3), abar_Z (0), abar_G (1), direct s bar
template <class CS, int ACT>
struct DgmRev1Fn {
  const float* A4; const float* S; const float* SBn; float* AB4; float* SBp; int Hp;
  DGMK_HD void operator()(int64_t i) const {
    DGMK_SMEM(A4); DGMK_SMEM(S); DGMK_SMEM(SBn); DGMK_SMEM(AB4); DGMK_SMEM(SBp);
    int64_t p = idiv(i, Hp); int j = (int)(i - p * Hp);
    const int64_t ld = 4 * (int64_t)Hp;
    const float* row = A4 + (p * CS::C) * ld + j;
    float* orow = AB4 + (p * CS::C) * ld + j;
    int64_t sb = (p * CS::C) * Hp + j;
    float afz[CS::C], afg[CS::C], afh[CS::C], z[CS::C], omg[CS::C], h[CS::C], s[CS::C], nb[CS::C];
#pragma unroll
    for (int c = 0; c < CS::C; ++c) {
      afz[c] = row[c * ld]; afg[c] = row[c * ld + Hp]; afh[c] = row[c * ld + 3 * Hp];
      s[c] = S[sb + (int64_t)c * Hp]; nb[c] = SBn[sb + (int64_t)c * Hp];
    }
    aform_to_jet<CS, ACT>(afz, z);
    aform_to_jet<CS, ACT>(afg, omg);
    aform_to_jet<CS, ACT>(afh, h);
#pragma unroll
    for (int c = 0; c < CS::C; ++c) omg[c] = -omg[c];
    omg[0] += 1.0f;
    float yb[CS::C], ab[CS::C];
    prod_adj<CS, false>(nb, omg, yb);  // Hbar
    act_adj<CS, ACT>(yb, afh, ab);
#pragma unroll
    for (int c = 0; c < CS::C; ++c) orow[c * ld + 3 * Hp] = ab[c];
    prod_adj<CS, false>(nb, h, yb);    // -(Gbar)
#pragma unroll
    for (int c = 0; c < CS::C; ++c) yb[c] = -yb[c];
    act_adj<CS, ACT>(yb, afg, ab);
#pragma unroll
    for (int c = 0; c < CS::C; ++c) orow[c * ld + Hp] = ab[c];
    prod_adj<CS, false>(nb, s, yb);    // Zbar
    act_adj<CS, ACT>(yb, afz, ab);
#pragma unroll
    for (int c = 0; c < CS::C; ++c) orow[c * ld] = ab[c];
    prod_adj<CS, false>(nb, z, yb);    // direct path to s
#pragma unroll
    for (int c = 0; c < CS::C; ++c) SBp[sb + (int64_t)c * Hp] = yb[c];
  }
  // the same arithmetic on values already in registers (runv below)
  DGMK_HD static void core(const float* afz, const float* afg, const float* afh, const float* s, const float* nb,
                           float* abz, float* abg, float* abh, float* sbp) {
    float z[CS::C], omg[CS::C], h[CS::C], yb[CS::C];
    aform_to_jet<CS, ACT>(afz, z);
    aform_to_jet<CS, ACT>(afg, omg);
    aform_to_jet<CS, ACT>(afh, h);
#pragma unroll
    for (int c = 0; c < CS::C; ++c) omg[c] = -omg[c];
    omg[0] += 1.0f;
    prod_adj<CS, false>(nb, omg, yb);  // Hbar
    act_adj<CS, ACT>(yb, afh, abh);
    prod_adj<CS, false>(nb, h, yb);    // -(Gbar)
#pragma unroll
    for (int c = 0; c < CS::C; ++c) yb[c] = -yb[c];
    act_adj<CS, ACT>(yb, afg, abg);
    prod_adj<CS, false>(nb, s, yb);    // Zbar
    act_adj<CS, ACT>(yb, afz, abz);
    prod_adj<CS, false>(nb, z, sbp);   // direct path to s
  }
  // V = 2 or 4 consecutive units of one point (Hp % V == 0): 8- / 16-byte accesses, the index division once
  // per V elements; k = p * (Hp/V) + j/V.  sink(u, slot, ab) is called with each pre-activation cotangent
  // (slot 3 = H, 1 = G, 0 = Z; u = unit within the group): the CUDA backend forms grad[U | b] from them
  template <int V> struct alignas(4 * V) Vec { float v[V]; };
  template <int V, class Sink>
  DGMK_HD void runv(int64_t k, Sink&& sink) const {
    const int q = Hp / V;
    int64_t p = idiv(k, q); int j = (int)(k - p * q) * V;
    const int64_t ld = 4 * (int64_t)Hp;
    const float* row = A4 + (p * CS::C) * ld + j;
    float* orow = AB4 + (p * CS::C) * ld + j;
    const int64_t sb = (p * CS::C) * Hp + j;
    Vec<V> vz[CS::C], vg[CS::C], vh[CS::C], vs[CS::C], vn[CS::C];
#pragma unroll
    for (int c = 0; c < CS::C; ++c) {
      vz[c] = *reinterpret_cast<const Vec<V>*>(row + c * ld); vg[c] = *reinterpret_cast<const Vec<V>*>(row + c * ld + Hp);
      vh[c] = *reinterpret_cast<const Vec<V>*>(row + c * ld + 3 * Hp);
      vs[c] = *reinterpret_cast<const Vec<V>*>(S + sb + (int64_t)c * Hp); vn[c] = *reinterpret_cast<const Vec<V>*>(SBn + sb + (int64_t)c * Hp);
    }
    Vec<V> oz[CS::C], og[CS::C], oh[CS::C], os[CS::C];
#pragma unroll
    for (int u = 0; u < V; ++u) {
      float afz[CS::C], afg[CS::C], afh[CS::C], s[CS::C], nb[CS::C], abz[CS::C], abg[CS::C], abh[CS::C], sbp[CS::C];
#pragma unroll
      for (int c = 0; c < CS::C; ++c) { afz[c] = vz[c].v[u]; afg[c] = vg[c].v[u]; afh[c] = vh[c].v[u]; s[c] = vs[c].v[u]; nb[c] = vn[c].v[u]; }
      core(afz, afg, afh, s, nb, abz, abg, abh, sbp);
      sink(u, 3, abh); sink(u, 1, abg); sink(u, 0, abz);
#pragma unroll
      for (int c = 0; c < CS::C; ++c) { oz[c].v[u] = abz[c]; og[c].v[u] = abg[c]; oh[c].v[u] = abh[c]; os[c].v[u] = sbp[c]; }
    }
#pragma unroll
    for (int c = 0; c < CS::C; ++c) {
      *reinterpret_cast<Vec<V>*>(orow + c * ld + 3 * Hp) = oh[c];
      *reinterpret_cast<Vec<V>*>(orow + c * ld + Hp) = og[c];
      *reinterpret_cast<Vec<V>*>(orow + c * ld) = oz[c];
      *reinterpret_cast<Vec<V>*>(SBp + sb + (int64_t)c * Hp) = os[c];
    }
  }
};
// reverse stage 2: (s*R)bar -> abar_R (slot 2), s bar += R * (sR)bar
template <class CS, int ACT>
struct DgmRev2Fn {
  const float* A4; const float* S; const float* SRB; float* AB4; float* SBp; int Hp;
  DGMK_HD void operator()(int64_t i) const {
    DGMK_SMEM(A4); DGMK_SMEM(S); DGMK_SMEM(SRB); DGMK_SMEM(AB4); DGMK_SMEM(SBp);
    int64_t p = idiv(i, Hp); int j = (int)(i - p * Hp);
    const int64_t ld = 4 * (int64_t)Hp;
    const float* row = A4 + (p * CS::C) * ld + 2 * Hp + j;
    float* orow = AB4 + (p * CS::C) * ld + 2 * Hp + j;
    int64_t sb = (p * CS::C) * Hp + j;
    float afr[CS::C], r[CS::C], s[CS::C], srb[CS::C], sbar[CS::C], yb[CS::C], ab[CS::C];
#pragma unroll
    for (int c = 0; c < CS::C; ++c) {
      afr[c] = row[c * ld]; s[c] = S[sb + (int64_t)c * Hp];
      srb[c] = SRB[sb + (int64_t)c * Hp]; sbar[c] = SBp[sb + (int64_t)c * Hp];
    }
    aform_to_jet<CS, ACT>(afr, r);
    prod_adj<CS, false>(srb, s, yb);  // Rbar
    act_adj<CS, ACT>(yb, afr, ab);
#pragma unroll
    for (int c = 0; c < CS::C; ++c) orow[c * ld] = ab[c];
    prod_adj<CS, true>(srb, r, sbar);
#pragma unroll
    for (int c = 0; c < CS::C; ++c) SBp[sb + (int64_t)c * Hp] = sbar[c];
  }
};

// ---- output layer reverse: s bar[r][j] = sum_m ubar[r][m] W_out[m][j] -----------
struct OutRevFn {
  const float* UB; const float* outw; float* SB; int Hp; int o;
  DGMK_HD void operator()(int64_t i) const {
    DGMK_SMEM(UB); DGMK_SMEM(SB);
    int64_t r = idiv(i, Hp); int j = (int)(i - r * Hp);
    float v = 0.f;
    for (int m = 0; m < o; ++m) v += UB[r * 4 + m] * outw[m * Hp + j];
    SB[i] = v;
  }
  DGMK_HD void vec4(int64_t k) const {   // four consecutive units of row r, same sums
    DGMK_SMEM(UB); DGMK_SMEM(SB);
    const int q = Hp >> 2;
    int64_t r = idiv(k, q); int j = (int)(k - r * q) * 4;
    const F4 ub = *reinterpret_cast<const F4*>(UB + r * 4);
    F4 v; v.x = 0.f; v.y = 0.f; v.z = 0.f; v.w = 0.f;
    for (int m = 0; m < o; ++m) {
      const float um = m == 0 ? ub.x : (m == 1 ? ub.y : (m == 2 ? ub.z : ub.w));
      const F4 w = *reinterpret_cast<const F4*>(outw + m * Hp + j);
      v.x += um * w.x; v.y += um * w.y; v.z += um * w.z; v.w += um * w.w;
    }
    *reinterpret_cast<F4*>(SB + r * (int64_t)Hp + j) = v;
  }
};

// OutRevFn followed by MlpRevFn of the top hidden layer in one pass: the output-layer cotangent y bar = UB outw is
// formed in registers (same sums as OutRevFn) instead of being written and re-read
template <class CS, int ACT>
struct OutMlpRevFn {
  const float* UB; const float* outw; const float* G; float* AB; int Hp; int o;
  DGMK_HD void operator()(int64_t i) const {
    DGMK_SMEM(UB); DGMK_SMEM(G); DGMK_SMEM(AB);
    int64_t p = idiv(i, Hp); int j = (int)(i - p * Hp);
    float af[CS::C], yb[CS::C], ab[CS::C];
    int64_t base = (p * CS::C) * Hp + j;
#pragma unroll
    for (int c = 0; c < CS::C; ++c) {
      af[c] = G[base + (int64_t)c * Hp];
      float v = 0.f;
      for (int m = 0; m < o; ++m) v += UB[(p * CS::C + c) * 4 + m] * outw[m * Hp + j];
      yb[c] = v;
    }
    act_adj<CS, ACT>(yb, af, ab);
#pragma unroll
    for (int c = 0; c < CS::C; ++c) AB[base + (int64_t)c * Hp] = ab[c];
  }
};

// ================================ losses =========================================
// U / UB are [rows*C][4]: output jets (column m = output component) and their
// cotangent seeds.  Lp[p] is the point's (already scaled) loss contribution.

// heat interior (heat.py:71-87): r = u_t - kappa u_xx
struct HeatInteriorFn {
  const float* U; float* UB; float* Lp; float kappa, inv;
  DGMK_HD void operator()(int64_t p) const {
    DGMK_SMEM(U); DGMK_SMEM(UB); DGMK_SMEM(Lp);
    const float* u = U + p * 16;
    float r = u[8] - kappa * u[12];
    float* ub = UB + p * 16;
    for (int q = 0; q < 16; ++q) ub[q] = 0.f;
    ub[8] = 2.f * r * inv;
    ub[12] = -2.f * kappa * r * inv;
    Lp[p] = r * r * inv;
  }
};
// value rows with targets (heat.py:89-94 IC/BC; simple_ode.py:62; fitzhugh_nagumo.py:95)
// per block b of the pass: mode 0 = target array tgt[b][row*o + m], 1 = sin(x[row][0])
struct ValueTargetFn {
  const float* U; float* UB; float* Lp; XSrc xs; const float* tgt[3]; int32_t mode[3];
  int32_t o; float inv;
  DGMK_HD void operator()(int64_t p) const {
    DGMK_SMEM(U); DGMK_SMEM(UB); DGMK_SMEM(Lp);
    int64_t b = (xs.block_rows >> 31) == 0 ? idiv(p, (int32_t)xs.block_rows) : p / xs.block_rows, w = p - b * xs.block_rows;
    float l = 0.f;
    for (int m = 0; m < 4; ++m) {
      float ubv = 0.f;
      if (m < o) {
        const int md = b == 0 ? mode[0] : (b == 1 ? mode[1] : mode[2]);
        const float* tg = b == 0 ? tgt[0] : (b == 1 ? tgt[1] : tgt[2]);
        float t = (md == 1) ? sinf(xs.at(p)[0]) : tg[w * o + m];
        float e = U[p * 4 + m] - t;
        l += e * e;
        ubv = 2.f * e * inv;
      }
      UB[p * 4 + m] = ubv;
    }
    Lp[p] = l * inv;
  }
};
// simple ODE interior (simple_ode.py:54-60): r = y_t + y
struct OdeInteriorFn {
  const float* U; float* UB; float* Lp; float inv;
  DGMK_HD void operator()(int64_t p) const {
    DGMK_SMEM(U); DGMK_SMEM(UB); DGMK_SMEM(Lp);
    const float* u = U + p * 8;
    float r = u[4] + u[0];
    float* ub = UB + p * 8;
    for (int q = 0; q < 8; ++q) ub[q] = 0.f;
    ub[0] = 2.f * r * inv;
    ub[4] = 2.f * r * inv;
    Lp[p] = r * r * inv;
  }
};
// FitzHugh-Nagumo interior (fitzhugh_nagumo.py:69-94)
struct FhnInteriorFn {
  const float* U; float* UB; float* Lp; float I, alpha, beta, tau, inv;
  DGMK_HD void operator()(int64_t p) const {
    DGMK_SMEM(U); DGMK_SMEM(UB); DGMK_SMEM(Lp);
    const float* u = U + p * 8;
    float Y = u[0], W = u[1], dY = u[4], dW = u[5];
    float rx = dY + (Y * Y * Y / 3.0f + W - I - Y);
    float ry = dW + (beta * W - alpha - Y) / tau;
    float* ub = UB + p * 8;
    for (int q = 0; q < 8; ++q) ub[q] = 0.f;
    float gx = 2.f * rx * inv, gy = 2.f * ry * inv;
    ub[4] = gx;
    ub[5] = gy;
    ub[0] = gx * (Y * Y - 1.f) - gy / tau;
    ub[1] = gx + gy * beta / tau;
    Lp[p] = (rx * rx + ry * ry) * inv;
  }
};
// Fredholm (fredholm.py:64-74): value rows of the chunk's points (Ux) and of their
// k Monte-Carlo nodes (Un, row = j*rows + p); accumulation order j = 0..k-1 as in
// the reference loop.
struct FredholmFn {
  const float* Ux; const float* Un; float* UBx; float* UBn; float* Lp;
  const float* x; const float* T; int64_t rows, Tstride; int32_t k; float dr, inv;
  DGMK_HD void operator()(int64_t p) const {
    float sx = sinf(x[p]);
    float integral = 0.f;
    for (int j = 0; j < k; ++j) {
      float w = sx * cosf(T[(int64_t)j * Tstride + p]);
      integral += w * Un[((int64_t)j * rows + p) * 4];
    }
    integral *= dr;
    float r = Ux[p * 4] - sx - integral;
    float g = 2.f * r * inv;
    UBx[p * 4] = g; UBx[p * 4 + 1] = 0.f; UBx[p * 4 + 2] = 0.f; UBx[p * 4 + 3] = 0.f;
    for (int j = 0; j < k; ++j) {
      float w = sx * cosf(T[(int64_t)j * Tstride + p]);
      float* ub = UBn + ((int64_t)j * rows + p) * 4;
      ub[0] = -g * dr * w; ub[1] = 0.f; ub[2] = 0.f; ub[3] = 0.f;
    }
    Lp[p] = r * r * inv;
  }
};

// The same loss in three pieces, for the resident-tile step (dgmk_tile.cuh), which walks the k nodes of a block of
// points in sub-tiles that fit shared memory: FredAccFn adds one sub-tile's terms to the running integral (same
// j = 0..k-1 order and FP32 arithmetic as FredholmFn), FredResFn forms the residual, the loss and the x-row seed and
// leaves g = 2 r / B in Ip, FredSeedFn seeds the node rows of one sub-tile.
struct FredAccFn {
  const float* Un; const float* x; const float* T; float* Ip; int64_t rows, Tstride; int32_t jj;
  DGMK_HD void operator()(int64_t p) const {
    DGMK_SMEM(Un); DGMK_SMEM(Ip);
    const float sx = sinf(x[p]);
    float integral = Ip[p];
    for (int j = 0; j < jj; ++j) {
      float w = sx * cosf(T[(int64_t)j * Tstride + p]);
      integral += w * Un[((int64_t)j * rows + p) * 4];
    }
    Ip[p] = integral;
  }
};
struct FredResFn {
  const float* Ux; float* UBx; float* Lp; const float* x; float* Ip; float dr, inv;
  DGMK_HD void operator()(int64_t p) const {
    DGMK_SMEM(Ux); DGMK_SMEM(UBx); DGMK_SMEM(Lp); DGMK_SMEM(Ip);
    const float sx = sinf(x[p]);
    float integral = Ip[p];
    integral *= dr;
    float r = Ux[p * 4] - sx - integral;
    float g = 2.f * r * inv;
    UBx[p * 4] = g; UBx[p * 4 + 1] = 0.f; UBx[p * 4 + 2] = 0.f; UBx[p * 4 + 3] = 0.f;
    Lp[p] = r * r * inv;
    Ip[p] = g;
  }
};
struct FredSeedFn {
  float* UBn; const float* x; const float* T; const float* G; int64_t rows, Tstride; float dr;
  DGMK_HD void operator()(int64_t i) const {
    DGMK_SMEM(UBn); DGMK_SMEM(G);
    int64_t j = idiv(i, (int32_t)rows), p = i - j * rows;
    float w = sinf(x[p]) * cosf(T[j * Tstride + p]);
    float g = G[p];
    float* ub = UBn + i * 4;
    ub[0] = -g * dr * w; ub[1] = 0.f; ub[2] = 0.f; ub[3] = 0.f;
  }
};

// ---- generic jet I/O for the module-level autograd seam -----------
// U -> Y[B,o], J[B,o,d], Hs[B,o,d,d]
template <class CS>
struct JetOutFn {
  const float* U; float* Y; float* J; float* Hs; int32_t o, d;
  DGMK_HD void operator()(int64_t p) const {
    const float* u = U + p * CS::C * 4;
    for (int m = 0; m < o; ++m) {
      Y[p * o + m] = u[m];
      if (J) for (int k = 0; k < CS::ND; ++k) J[(p * o + m) * d + k] = u[(1 + k) * 4 + m];
      if (Hs) {
#pragma unroll
        for (int q = 0; q < CS::NP; ++q) {
          float v = u[(1 + CS::ND + q) * 4 + m];
          Hs[((p * o + m) * d + CS::pi(q)) * d + CS::pj(q)] = v;
          Hs[((p * o + m) * d + CS::pj(q)) * d + CS::pi(q)] = v;
        }
      }
    }
  }
};
// cotangents (gY, gJ, gHs; any may be null) -> UB seeds
template <class CS>
struct JetSeedFn {
  float* UB; const float* gY; const float* gJ; const float* gHs; int32_t o, d;
  DGMK_HD void operator()(int64_t p) const {
    float* ub = UB + p * CS::C * 4;
    for (int m = 0; m < 4; ++m) {
      bool live = m < o;
      ub[m] = (live && gY) ? gY[p * o + m] : 0.f;
#pragma unroll
      for (int k = 0; k < CS::ND; ++k) ub[(1 + k) * 4 + m] = (live && gJ) ? gJ[(p * o + m) * d + k] : 0.f;
#pragma unroll
      for (int q = 0; q < CS::NP; ++q) {
        float v = 0.f;
        if (live && gHs) {
          v = gHs[((p * o + m) * d + CS::pi(q)) * d + CS::pj(q)];
          if (CS::pi(q) != CS::pj(q)) v += gHs[((p * o + m) * d + CS::pj(q)) * d + CS::pi(q)];
        }
        ub[(1 + CS::ND + q) * 4 + m] = v;
      }
    }
  }
};

// ---- fused Adam on the flat buffers (torch.optim.Adam defaults, heat.py:115) -----
struct AdamFn {
  float* theta; float* m; float* v; const float* g; const uint8_t* live;
  float step_size, bc2_sqrt, w1, b2, w2, eps;  // lr/(1-b1^t), sqrt(1-b2^t), 1-b1, b2, 1-b2
  DGMK_HD void operator()(int64_t i) const {
    if (live && !live[i]) return;  // grad is None -> torch skips the parameter
    float gi = g[i];
    float mi = m[i] + w1 * (gi - m[i]);      // exp_avg.lerp_(grad, 1 - beta1)
    float vi = b2 * v[i] + (w2 * gi) * gi;   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    theta[i] -= step_size * (mi / denom);    // param.addcdiv_(exp_avg, denom, value=-step_size)
  }
};

// ================================ on-device collocation sampler =====================
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11; Random123 philox4x32_R(10, ...)): counter-based, so a sample is a
// pure function of (seed, stream, step, element) -- no generator state to carry through a CUDA graph.  Replaces the
// torch.rand / rand_like draws of the reference drivers (heat.py:125-126, simple_ode.py:91, fitzhugh_nagumo.py:129,
// fredholm.py:67,100) when the driver is asked for sampler="philox"; bit-exact oracle: oracle/philox_np.py.
DGMK_HD void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    c[1] = (uint32_t)p1; c[3] = (uint32_t)p0; c[0] = n0; c[2] = n2;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
// u in [0, 1) with 24 random bits; value = fl(fl(span * u) + lo) (two roundings: no FMA contraction, so that the
// numpy oracle reproduces every bit)
DGMK_HD float philox_value(uint32_t w, float lo, float span) {
  const float u = (float)(w >> 8) * 5.9604644775390625e-8f;
#if defined(__CUDA_ARCH__)
  return __fadd_rn(__fmul_rn(span, u), lo);
#else
  volatile float t = span * u;
  return t + lo;
#endif
}
// block b of stream `stream_id` at step (*step_dev + step_add): words for elements 4b .. 4b+3
struct PhiloxKey {
  uint32_t k0, k1; const long long* step_dev; long long step_add;
  DGMK_HD void block(int64_t b, uint32_t stream_id, uint32_t (&c)[4]) const {
    const long long step = step_add + (step_dev ? *step_dev : 0);
    c[0] = (uint32_t)b; c[1] = (uint32_t)((uint64_t)b >> 32); c[2] = (uint32_t)step; c[3] = stream_id;
    philox4x32_10(c, k0, k1);
  }
};
// out[i] = lo + span * u_i; one item = four consecutive elements
struct PhiloxUniformFn {
  float* out; int64_t n; float lo, span; uint32_t stream_id; PhiloxKey key;
  DGMK_HD void operator()(int64_t b) const {
    uint32_t c[4];
    key.block(b, stream_id, c);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (4 * b + q < n) out[4 * b + q] = philox_value(c[q], lo, span);
  }
};
// the four operand blocks of a heat step (heat.py:125-134): x = xmax u (stream 0), t = tmax u' (stream 1);
// X = [x, t], X0 = [x, 0], XBD1 = [0, t], XBD2 = [xbd2, t]; one item = four consecutive points
struct PhiloxHeatFn {
  float* X; float* X0; float* XBD1; float* XBD2; int64_t B; float xmax, tmax, xbd2; PhiloxKey key;
  DGMK_HD void operator()(int64_t b) const {
    uint32_t cx[4], ct[4];
    key.block(b, 0u, cx);
    key.block(b, 1u, ct);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t i = 4 * b + q;
      if (i < B) {
        const float x = philox_value(cx[q], 0.f, xmax), t = philox_value(ct[q], 0.f, tmax);
        X[2 * i] = x; X[2 * i + 1] = t;
        X0[2 * i] = x; X0[2 * i + 1] = 0.f;
        XBD1[2 * i] = 0.f; XBD1[2 * i + 1] = t;
        XBD2[2 * i] = xbd2; XBD2[2 * i + 1] = t;
      }
    }
  }
};

// Device-resident step counter variant (CUDA-graph capturable training loops): state = 16 bytes,
// [int64 step | float lr/(1-b1^t) | float sqrt(1-b2^t)].  AdamPrepFn (one thread) advances the
// counter and forms the two bias-correction scalars in doubles exactly like torch does on the host;
// AdamDevFn is AdamFn with those two scalars read from the state.
struct AdamPrepFn {
  long long* state; double lr, b1, b2;
  DGMK_HD void operator()(int64_t) const {
    const long long t = state[0] + 1;
    state[0] = t;
    float* sc = reinterpret_cast<float*>(state + 1);
    sc[0] = (float)(lr / (1.0 - pow(b1, (double)t)));
    sc[1] = (float)sqrt(1.0 - pow(b2, (double)t));
  }
};
struct AdamDevFn {
  float* theta; float* m; float* v; const float* g; const uint8_t* live; const long long* state;
  float w1, b2, w2, eps;
  DGMK_HD void operator()(int64_t i) const {
    if (live && !live[i]) return;
    const float* sc = reinterpret_cast<const float*>(state + 1);
    const float step_size = sc[0], bc2_sqrt = sc[1];
    float gi = g[i];
    float mi = m[i] + w1 * (gi - m[i]);
    float vi = b2 * v[i] + (w2 * gi) * gi;
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    theta[i] -= step_size * (mi / denom);
  }
};

}  // namespace dgmk
