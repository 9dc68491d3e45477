// Implementation of the C ABI in include/dgmk.h, templated on the backend.
// dgmk_cuda.cu instantiates it with the CUDA backend (the product); the test-only
// host harness instantiates it with loops.  See dgmk_pipeline.h.
#pragma once
#include <math.h>
#include <stdio.h>
#include "../../include/dgmk.h"
#include "dgmk_pipeline.h"
#include "dgmk_steps.h"

namespace dgmk {

static thread_local char g_err[512] = "";
inline int fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

inline bool dims_from_desc(const dgmk_net_desc* d, NetDims* n) {
  if (!d) { fail(DGMK_EINVAL, "desc is NULL"); return false; }
  const char* err = "";
  if (!make_dims(d->kind, d->input_dim, d->output_dim, d->hidden_size, d->num_layers, d->activation, n, &err)) {
    fail(DGMK_EINVAL, err);
    return false;
  }
  return true;
}

inline bool init_ctx(const dgmk_net_desc* d, Ctx* c) {
  if (!dims_from_desc(d, &c->n)) return false;
  make_packed_layout(c->n, &c->pl);
  const char* err = "";
  if (!make_seg_tables(c->n, c->pl, &c->pack, &c->grad, &err)) { fail(DGMK_EINVAL, err); return false; }
  return true;
}

// bytes of the per-chunk region for `ch` points of a given problem class
inline size_t chunk_region_bytes(const NetDims& n, int cls, int64_t ch, int k, bool inplace = false) {
  switch (cls) {
    case DGMK_WS_HEAT: {
      size_t a = pass_bytes(n, ch, CS_HEAT) + rev_bytes(n, 4 * ch, inplace);
      size_t b = pass_bytes(n, 3 * ch, CS_V) + rev_bytes(n, 3 * ch, inplace);
      return a > b ? a : b;
    }
    case DGMK_WS_ODE:
    case DGMK_WS_FHN: {
      size_t a = pass_bytes(n, ch, CS_D1O1) + rev_bytes(n, 2 * ch, inplace);
      size_t b = pass_bytes(n, ch, CS_V) + rev_bytes(n, ch, inplace);
      return a > b ? a : b;
    }
    case DGMK_WS_FREDHOLM:
      return pass_bytes(n, ch, CS_V) + pass_bytes(n, ch * k, CS_V) + rev_bytes(n, ch * k);
    case DGMK_WS_JET0: return pass_bytes(n, ch, CS_V) + rev_bytes(n, ch);
    case DGMK_WS_JET1: {
      int cs = n.d == 1 ? CS_D1O1 : CS_D2O1;
      return pass_bytes(n, ch, cs) + rev_bytes(n, ch * cs_channels(cs));
    }
    default: {
      int cs = n.d == 1 ? CS_D1O2 : CS_D2O2;
      return pass_bytes(n, ch, cs) + rev_bytes(n, ch * cs_channels(cs));
    }
  }
}
inline int64_t loss_points(int cls, int64_t ch, int k) {
  (void)k;
  return cls == DGMK_WS_HEAT ? 3 * ch : ch;
}
inline size_t total_bytes(const Ctx& c, int cls, int64_t ch, int k) {
  return ctx_bytes(c.n, c.pl, loss_points(cls, ch, k)) + chunk_region_bytes(c.n, cls, ch, k) + 4096;
}
constexpr size_t WS_TARGET = (size_t)40 << 30;  // recommended workspace cap (B200: 180 GB HBM3e)
inline int64_t recommended_chunk(const Ctx& c, int cls, int64_t B, int k) {
  if (B <= 1024) return B;
  if (total_bytes(c, cls, B, k) <= WS_TARGET) return B;
  int64_t lo = 1024, hi = B;  // largest multiple of 1024 under the target
  while (hi - lo > 1024) {
    int64_t mid = (lo + hi) / 2 / 1024 * 1024;
    if (mid <= lo) break;
    if (total_bytes(c, cls, mid, k) <= WS_TARGET) lo = mid; else hi = mid;
  }
  return lo;
}
// largest chunk (<= B) that fits the workspace actually given
inline int64_t fit_chunk(const Ctx& c, int cls, int64_t B, int k, size_t ws_bytes) {
  if (total_bytes(c, cls, B, k) <= ws_bytes) return B;
  int64_t lo = 0, hi = B;
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) / 2;
    if (total_bytes(c, cls, mid, k) <= ws_bytes) lo = mid; else hi = mid;
  }
  if (lo >= 1024) lo = lo / 1024 * 1024;
  return lo;
}

template <class BK>
struct Api {
  static int check_common(BK& bk, const void* const* dev_ptrs, int nptr, int64_t B, int64_t Bg) {
    if (B <= 0 || Bg < B) return fail(DGMK_EINVAL, "need 0 < B <= B_global");
    for (int i = 0; i < nptr; ++i) {
      if (!dev_ptrs[i]) return fail(DGMK_EINVAL, "NULL buffer");
      if (!bk.is_device_ptr(dev_ptrs[i])) return fail(DGMK_EDEVICE, "host pointer passed: dgmk has no CPU path, all buffers must be device memory");
    }
    return 0;
  }
  static int finish(BK& bk) {
    const char* e = bk.error();
    if (e) return fail(DGMK_ECUDA, e);
    return 0;
  }

  // ------------------------------------------------------------------ heat
  static int heat_step(const dgmk_net_desc* desc, const float* theta, const float* x, const float* x0,
                       const float* xbd1, const float* xbd2, const float* t_bd1, const float* t_bd2,
                       int64_t B, int64_t Bg, float kappa, float* loss, float* grad, void* ws, size_t wsb, void* stream) {
    Ctx c; if (!init_ctx(desc, &c)) return DGMK_EINVAL;
    if (c.n.d != 2 || c.n.o != 1) return fail(DGMK_EINVAL, "heat step needs input_dim=2, output_dim=1");
    BK bk(stream);
    const void* ptrs[] = {theta, x, x0, xbd1, xbd2, t_bd1, t_bd2, loss, grad, ws};
    if (int r = check_common(bk, ptrs, 10, B, Bg)) return r;
    int64_t ch = fit_chunk(c, DGMK_WS_HEAT, B, 0, wsb);
    if (ch < 1) return fail(DGMK_EWORKSPACE, "workspace too small (see dgmk_workspace_bytes)");
    Carver cv(ws, wsb);
    if (!carve_ctx(cv, &c, 3 * ch)) return fail(DGMK_EWORKSPACE, "workspace too small");
    Pipeline<BK> P(bk, c);
    P.pack(theta);
    HeatArgs ha; ha.x = x; ha.x0 = x0; ha.xbd1 = xbd1; ha.xbd2 = xbd2; ha.t_bd1 = t_bd1; ha.t_bd2 = t_bd2;
    ha.kappa = kappa; ha.inv = (float)(1.0 / (double)Bg);
    if constexpr (BK::kHasTile) {   // hidden sizes <= 64: the whole step in one persistent kernel (dgmk_tile.cuh)
      bk.set_unpack_target(grad, loss);
      if (bk.tile_step(c, DGMK_WS_HEAT, &ha, nullptr, B)) { if (!bk.unpacked) P.unpack(grad, loss); return finish(bk); }
    }
    P.zero_grads();
    const size_t mark = cv.off;
    for (int64_t p0 = 0; p0 < B; p0 += ch) {
      int64_t r = (B - p0 < ch) ? B - p0 : ch;
      if (!heat_chunk(P, cv, mark, ha, p0, r)) return fail(DGMK_EWORKSPACE, "workspace too small");
    }
    P.unpack(grad, loss);
    return finish(bk);
  }

  // ------------------------------------------------------------------ ode / fhn
  static int ode_like_step(bool fhn, const dgmk_net_desc* desc, const float* theta, const float* t, const float* t0,
                           const float* y_ic, int64_t B, int64_t Bg, float* loss, float* grad, void* ws, size_t wsb,
                           void* stream) {
    Ctx c; if (!init_ctx(desc, &c)) return DGMK_EINVAL;
    if (c.n.d != 1 || c.n.o != (fhn ? 2 : 1)) return fail(DGMK_EINVAL, fhn ? "fhn step needs input_dim=1, output_dim=2" : "ode step needs input_dim=1, output_dim=1");
    BK bk(stream);
    const void* ptrs[] = {theta, t, t0, y_ic, loss, grad, ws};
    if (int r = check_common(bk, ptrs, 7, B, Bg)) return r;
    const int cls = fhn ? DGMK_WS_FHN : DGMK_WS_ODE;
    int64_t ch = fit_chunk(c, cls, B, 0, wsb);
    if (ch < 1) return fail(DGMK_EWORKSPACE, "workspace too small (see dgmk_workspace_bytes)");
    Carver cv(ws, wsb);
    if (!carve_ctx(cv, &c, ch)) return fail(DGMK_EWORKSPACE, "workspace too small");
    Pipeline<BK> P(bk, c);
    P.pack(theta);
    OdeArgs oa; oa.t = t; oa.t0 = t0; oa.y_ic = y_ic; oa.inv = (float)(1.0 / (double)Bg); oa.fhn = fhn ? 1 : 0;
    if constexpr (BK::kHasTile) {
      bk.set_unpack_target(grad, loss);
      if (bk.tile_step(c, cls, nullptr, &oa, B)) { if (!bk.unpacked) P.unpack(grad, loss); return finish(bk); }
    }
    P.zero_grads();
    const size_t mark = cv.off;
    for (int64_t p0 = 0; p0 < B; p0 += ch) {
      int64_t r = (B - p0 < ch) ? B - p0 : ch;
      if (!ode_like_chunk(P, cv, mark, oa, p0, r)) return fail(DGMK_EWORKSPACE, "workspace too small");
    }
    P.unpack(grad, loss);
    return finish(bk);
  }

  // ------------------------------------------------------------------ fredholm
  static int fredholm_step(const dgmk_net_desc* desc, const float* theta, const float* x, const float* nodes, int64_t B,
                           int32_t k, int64_t Bg, float* loss, float* grad, void* ws, size_t wsb, void* stream) {
    Ctx c; if (!init_ctx(desc, &c)) return DGMK_EINVAL;
    if (c.n.d != 1 || c.n.o != 1) return fail(DGMK_EINVAL, "fredholm step needs input_dim=1, output_dim=1");
    if (k < 1) return fail(DGMK_EINVAL, "k must be >= 1");
    BK bk(stream);
    const void* ptrs[] = {theta, x, nodes, loss, grad, ws};
    if (int r = check_common(bk, ptrs, 6, B, Bg)) return r;
    int64_t ch = fit_chunk(c, DGMK_WS_FREDHOLM, B, k, wsb);
    if (ch < 1) return fail(DGMK_EWORKSPACE, "workspace too small (see dgmk_workspace_bytes)");
    Carver cv(ws, wsb);
    if (!carve_ctx(cv, &c, ch)) return fail(DGMK_EWORKSPACE, "workspace too small");
    Pipeline<BK> P(bk, c);
    P.pack(theta);
    const float inv = (float)(1.0 / (double)Bg);
    const float dr = (float)(M_PI / (2.0 * k));
    FredArgs fa; fa.x = x; fa.nodes = nodes; fa.B = B; fa.k = k; fa.dr = dr; fa.inv = inv;
    if constexpr (BK::kHasTile) {   // hidden sizes <= 64: blocks of points, node sub-tiles in shared memory (dgmk_tile.cuh)
      bk.set_unpack_target(grad, loss);
      if (bk.tile_step_fredholm(c, fa)) { if (!bk.unpacked) P.unpack(grad, loss); return finish(bk); }
    }
    P.zero_grads();
    if (bk.fredholm_block_nodes() > 0) {   // (test harness) the sub-tiled body on the host
      const int64_t rb_ = bk.fredholm_block_points() < ch ? bk.fredholm_block_points() : ch;
      float* Ip = cv.take(rb_);
      if (!Ip) return fail(DGMK_EWORKSPACE, "workspace too small");
      const size_t markb = cv.off;
      for (int64_t p0 = 0; p0 < B; p0 += rb_) {
        int64_t r = (B - p0 < rb_) ? B - p0 : rb_;
        if (!fredholm_block(P, cv, markb, fa, p0, r, bk.fredholm_block_nodes(), Ip)) return fail(DGMK_EWORKSPACE, "workspace too small");
      }
      P.unpack(grad, loss);
      return finish(bk);
    }
    const size_t mark = cv.off;
    for (int64_t p0 = 0; p0 < B; p0 += ch) {
      int64_t r = (B - p0 < ch) ? B - p0 : ch;
      cv.off = mark;
      PassBufs px, pn; RevBufs rb;
      px.xs = xsrc1(x + p0, r, 1);
      pn.xs = xsrc1(nodes + p0, r, 1);
      pn.xs.block_stride = B;  // nodes[j] starts B floats after nodes[j-1]
      if (!carve_pass(cv, c.n, &px, r, CS_V) || !carve_pass(cv, c.n, &pn, r * k, CS_V) || !carve_rev(cv, c.n, &rb, pn.M))
        return fail(DGMK_EWORKSPACE, "workspace too small");
      P.forward(px);
      P.forward(pn);
      FredholmFn f; f.Ux = px.U; f.Un = pn.U; f.UBx = px.UB; f.UBn = pn.UB; f.Lp = c.Lp;
      f.x = x + p0; f.T = nodes + p0; f.rows = r; f.Tstride = B; f.k = k; f.dr = dr; f.inv = inv;
      bk.ew(f, r);
      P.add_loss(r);
      P.reverse(px, rb);
      P.reverse(pn, rb);
    }
    P.unpack(grad, loss);
    return finish(bk);
  }

  // ------------------------------------------------------------------ jets
  static int jet_cs(const NetDims& n, int order) {
    if (order == 0) return CS_V;
    if (order == 1) return n.d == 1 ? CS_D1O1 : CS_D2O1;
    return n.d == 1 ? CS_D1O2 : CS_D2O2;
  }
  static int jet_forward(const dgmk_net_desc* desc, const float* theta, const float* x, int64_t B, int order, float* Y,
                         float* J, float* Hs, void* ws, size_t wsb, void* stream) {
    Ctx c; if (!init_ctx(desc, &c)) return DGMK_EINVAL;
    if (order < 0 || order > 2) return fail(DGMK_EINVAL, "order must be 0, 1 or 2");
    BK bk(stream);
    const void* ptrs[] = {theta, x, Y, ws};
    if (int r = check_common(bk, ptrs, 4, B, B)) return r;
    if ((order >= 1 && !J) || (order >= 2 && !Hs)) return fail(DGMK_EINVAL, "J / Hs output buffer missing");
    const int cls = DGMK_WS_JET0 + order;
    if (total_bytes(c, cls, B, 0) > wsb) return fail(DGMK_EWORKSPACE, "jet workspace must hold the whole batch (dgmk_workspace_bytes)");
    Carver cv(ws, wsb);
    carve_ctx(cv, &c, B);
    Pipeline<BK> P(bk, c);
    P.pack(theta);
    PassBufs pb;
    pb.xs = xsrc1(x, B, c.n.d);
    const int cs = jet_cs(c.n, order);
    if (!carve_pass(cv, c.n, &pb, B, cs)) return fail(DGMK_EWORKSPACE, "workspace too small");
    P.forward(pb);
    DGMK_CS_SWITCH(cs, CS, {
      JetOutFn<CS> f; f.U = pb.U; f.Y = Y; f.J = order >= 1 ? J : nullptr; f.Hs = order >= 2 ? Hs : nullptr; f.o = c.n.o; f.d = c.n.d;
      bk.ew(f, B);
    })
    return finish(bk);
  }
  static int jet_reverse(const dgmk_net_desc* desc, const float* theta, const float* x, int64_t B, int order,
                         const float* gY, const float* gJ, const float* gHs, float* grad, void* ws, size_t wsb, void* stream) {
    Ctx c; if (!init_ctx(desc, &c)) return DGMK_EINVAL;
    if (order < 0 || order > 2) return fail(DGMK_EINVAL, "order must be 0, 1 or 2");
    BK bk(stream);
    const void* ptrs[] = {theta, x, grad, ws};
    if (int r = check_common(bk, ptrs, 4, B, B)) return r;
    const int cls = DGMK_WS_JET0 + order;
    if (total_bytes(c, cls, B, 0) > wsb) return fail(DGMK_EWORKSPACE, "jet workspace must hold the whole batch (dgmk_workspace_bytes)");
    Carver cv(ws, wsb);
    carve_ctx(cv, &c, B);
    Pipeline<BK> P(bk, c);
    P.pack(theta);
    P.zero_grads();
    PassBufs pb; RevBufs rb;
    pb.xs = xsrc1(x, B, c.n.d);
    const int cs = jet_cs(c.n, order);
    if (!carve_pass(cv, c.n, &pb, B, cs) || !carve_rev(cv, c.n, &rb, pb.M)) return fail(DGMK_EWORKSPACE, "workspace too small");
    DGMK_CS_SWITCH(cs, CS, {
      JetSeedFn<CS> f; f.UB = pb.UB; f.gY = gY; f.gJ = order >= 1 ? gJ : nullptr; f.gHs = order >= 2 ? gHs : nullptr; f.o = c.n.o; f.d = c.n.d;
      bk.ew(f, B);
    })
    P.reverse(pb, rb);
    P.unpack(grad, nullptr);
    return finish(bk);
  }
  static int eval(const dgmk_net_desc* desc, const float* theta, const float* x, int64_t B, float* Y, void* ws, size_t wsb,
                  void* stream) {
    Ctx c; if (!init_ctx(desc, &c)) return DGMK_EINVAL;
    BK bk(stream);
    const void* ptrs[] = {theta, x, Y, ws};
    if (int r = check_common(bk, ptrs, 4, B, B)) return r;
    int64_t ch = fit_chunk(c, DGMK_WS_JET0, B, 0, wsb);
    if (ch < 1) return fail(DGMK_EWORKSPACE, "workspace too small (see dgmk_workspace_bytes)");
    Carver cv(ws, wsb);
    if (!carve_ctx(cv, &c, ch)) return fail(DGMK_EWORKSPACE, "workspace too small");
    Pipeline<BK> P(bk, c);
    P.pack(theta);
    const size_t mark = cv.off;
    for (int64_t p0 = 0; p0 < B; p0 += ch) {
      int64_t r = (B - p0 < ch) ? B - p0 : ch;
      cv.off = mark;
      PassBufs pb;
      pb.xs = xsrc1(x + p0 * c.n.d, r, c.n.d);
      if (!carve_pass(cv, c.n, &pb, r, CS_V)) return fail(DGMK_EWORKSPACE, "workspace too small");
      P.forward(pb);
      JetOutFn<CsV> f; f.U = pb.U; f.Y = Y + p0 * c.n.o; f.J = nullptr; f.Hs = nullptr; f.o = c.n.o; f.d = c.n.d;
      bk.ew(f, r);
    }
    return finish(bk);
  }
  static int adam(float* theta, float* m, float* v, const float* g, const uint8_t* live, int64_t P, double lr, double b1,
                  double b2, double eps, int64_t step, void* stream) {
    if (P <= 0 || step < 1) return fail(DGMK_EINVAL, "need P > 0 and step >= 1");
    BK bk(stream);
    const void* ptrs[] = {theta, m, v, g};
    if (int r = check_common(bk, ptrs, 4, 1, 1)) return r;
    if (live && !bk.is_device_ptr(live)) return fail(DGMK_EDEVICE, "live mask must be device memory");
    AdamFn f; f.theta = theta; f.m = m; f.v = v; f.g = g; f.live = live;
    // same scalar arithmetic as torch/optim/adam.py::_single_tensor_adam (doubles on the host)
    double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
    f.step_size = (float)(lr / bc1);
    f.bc2_sqrt = (float)sqrt(bc2);
    f.w1 = (float)(1.0 - b1);
    f.b2 = (float)b2; f.w2 = (float)(1.0 - b2); f.eps = (float)eps;
    bk.ew(f, P);
    return finish(bk);
  }
  // same update with the step counter (and the two bias-correction scalars) in device memory:
  // nothing in the launch depends on host state, so a captured CUDA graph can replay it
  static int adam_dev(float* theta, float* m, float* v, const float* g, const uint8_t* live, int64_t P, double lr,
                      double b1, double b2, double eps, long long* state, void* stream) {
    if (P <= 0 || !state) return fail(DGMK_EINVAL, "need P > 0 and a 16-byte device state");
    BK bk(stream);
    const void* ptrs[] = {theta, m, v, g, state};
    if (int r = check_common(bk, ptrs, 5, 1, 1)) return r;
    if (live && !bk.is_device_ptr(live)) return fail(DGMK_EDEVICE, "live mask must be device memory");
    AdamPrepFn pf; pf.state = state; pf.lr = lr; pf.b1 = b1; pf.b2 = b2;
    bk.ew(pf, 1);
    AdamDevFn f; f.theta = theta; f.m = m; f.v = v; f.g = g; f.live = live; f.state = state;
    f.w1 = (float)(1.0 - b1);
    f.b2 = (float)b2; f.w2 = (float)(1.0 - b2); f.eps = (float)eps;
    bk.ew(f, P);
    return finish(bk);
  }
  // ------------------------------------------------------------------ on-device sampler
  static PhiloxKey philox_key(unsigned long long seed, const long long* step_dev, long long step_add) {
    PhiloxKey k; k.k0 = (uint32_t)seed; k.k1 = (uint32_t)(seed >> 32); k.step_dev = step_dev; k.step_add = step_add;
    return k;
  }
  static int sample_uniform(float* out, int64_t n, float lo, float hi, unsigned long long seed, uint32_t stream_id,
                            const long long* step_dev, long long step_add, void* stream) {
    if (n <= 0) return fail(DGMK_EINVAL, "need n > 0");
    BK bk(stream);
    const void* ptrs[] = {out};
    if (int r = check_common(bk, ptrs, 1, 1, 1)) return r;
    if (step_dev && !bk.is_device_ptr(step_dev)) return fail(DGMK_EDEVICE, "step counter must be device memory");
    PhiloxUniformFn f; f.out = out; f.n = n; f.lo = lo; f.span = hi - lo; f.stream_id = stream_id;
    f.key = philox_key(seed, step_dev, step_add);
    bk.ew(f, (n + 3) / 4);
    return finish(bk);
  }
  static int sample_heat(float* X, float* X0, float* XBD1, float* XBD2, int64_t B, float xmax, float tmax, float xbd2,
                         unsigned long long seed, const long long* step_dev, long long step_add, void* stream) {
    if (B <= 0) return fail(DGMK_EINVAL, "need B > 0");
    BK bk(stream);
    const void* ptrs[] = {X, X0, XBD1, XBD2};
    if (int r = check_common(bk, ptrs, 4, 1, 1)) return r;
    if (step_dev && !bk.is_device_ptr(step_dev)) return fail(DGMK_EDEVICE, "step counter must be device memory");
    PhiloxHeatFn f; f.X = X; f.X0 = X0; f.XBD1 = XBD1; f.XBD2 = XBD2; f.B = B; f.xmax = xmax; f.tmax = tmax; f.xbd2 = xbd2;
    f.key = philox_key(seed, step_dev, step_add);
    bk.ew(f, (B + 3) / 4);
    return finish(bk);
  }
};

inline int param_layout(const dgmk_net_desc* desc, int32_t index, int64_t* offset, int32_t* rows, int32_t* cols, int32_t* live) {
  NetDims n; if (!dims_from_desc(desc, &n)) return DGMK_EINVAL;
  // enumerate tensors in named_parameters() order
  struct T { int rows, cols, live; };
  T t[16 * 13 + 8]; int nt = 0;
  auto add = [&](int r, int cc, int lv) { t[nt].rows = r; t[nt].cols = cc; t[nt].live = lv; ++nt; };
  const int H = n.H, d = n.d, o = n.o;
  add(H, d, 1); add(H, 0, 1);  // cols == 0 marks a 1-D tensor of `rows` elements
  if (n.kind == KIND_MLP) { for (int l = 0; l < n.L; ++l) { add(H, H, 1); add(H, 0, 1); } }
  else if (n.kind == KIND_DGM_LINEAR) { for (int l = 0; l < n.L; ++l) for (int g = 0; g < 4; ++g) { add(H, H, 1); add(H, 0, 1); add(H, d, 1); } }
  else {
    for (int l = -1; l < n.L; ++l) {
      for (int g = 0; g < 4; ++g) add(d, H, l >= 0);
      for (int g = 0; g < 4; ++g) add(H, H, l >= 0);
      for (int g = 0; g < 4; ++g) add(1, H, l >= 0);
    }
  }
  add(o, H, 1); add(o, 0, 1);
  if (index < 0) return nt;
  if (index >= nt) return fail(DGMK_EINVAL, "tensor index out of range");
  int64_t off = 0;
  for (int i = 0; i < index; ++i) off += (int64_t)t[i].rows * (t[i].cols ? t[i].cols : 1);
  if (offset) *offset = off;
  if (rows) *rows = t[index].rows;
  if (cols) *cols = t[index].cols;
  if (live) *live = t[index].live;
  return 0;
}

inline size_t workspace_bytes(const dgmk_net_desc* desc, int32_t cls, int64_t B, int32_t k) {
  Ctx c; if (!init_ctx(desc, &c)) return 0;
  if (cls < 0 || cls > DGMK_WS_JET2 || B < 1) { fail(DGMK_EINVAL, "bad workspace class or B"); return 0; }
  if (cls == DGMK_WS_FREDHOLM && k < 1) { fail(DGMK_EINVAL, "k must be >= 1"); return 0; }
  if (cls >= DGMK_WS_JET0) return total_bytes(c, cls, B, 0);  // stash must hold all rows
  int64_t ch = recommended_chunk(c, cls, B, k);
  return total_bytes(c, cls, ch, k);
}

}  // namespace dgmk

// Defines the extern "C" surface for backend type BK.
#define DGMK_DEFINE_C_API(BK, BACKEND_NAME)                                                                        \
  extern "C" {                                                                                                     \
  int dgmk_version(void) { return DGMK_VERSION; }                                                                  \
  const char* dgmk_backend(void) { return BACKEND_NAME; }                                                          \
  const char* dgmk_last_error(void) { return dgmk::g_err; }                                                        \
  int64_t dgmk_param_count(const dgmk_net_desc* d) {                                                               \
    dgmk::NetDims n; if (!dgmk::dims_from_desc(d, &n)) return -1; return dgmk::num_params(n); }                    \
  int dgmk_param_layout(const dgmk_net_desc* d, int32_t i, int64_t* off, int32_t* r, int32_t* c, int32_t* lv) {   \
    return dgmk::param_layout(d, i, off, r, c, lv); }                                                              \
  size_t dgmk_workspace_bytes(const dgmk_net_desc* d, int32_t cls, int64_t B, int32_t k) {                         \
    return dgmk::workspace_bytes(d, cls, B, k); }                                                                  \
  int dgmk_heat_step(const dgmk_net_desc* d, const float* th, const float* x, const float* x0, const float* b1,    \
                     const float* b2, const float* t1, const float* t2, int64_t B, int64_t Bg, float kappa,        \
                     float* loss, float* g, void* ws, size_t wsb, void* st) {                                      \
    return dgmk::Api<BK>::heat_step(d, th, x, x0, b1, b2, t1, t2, B, Bg, kappa, loss, g, ws, wsb, st); }           \
  int dgmk_ode_step(const dgmk_net_desc* d, const float* th, const float* t, const float* t0, const float* yic,    \
                    int64_t B, int64_t Bg, float* loss, float* g, void* ws, size_t wsb, void* st) {                \
    return dgmk::Api<BK>::ode_like_step(false, d, th, t, t0, yic, B, Bg, loss, g, ws, wsb, st); }                  \
  int dgmk_fhn_step(const dgmk_net_desc* d, const float* th, const float* t, const float* t0, const float* yic,    \
                    int64_t B, int64_t Bg, float* loss, float* g, void* ws, size_t wsb, void* st) {                \
    return dgmk::Api<BK>::ode_like_step(true, d, th, t, t0, yic, B, Bg, loss, g, ws, wsb, st); }                   \
  int dgmk_fredholm_step(const dgmk_net_desc* d, const float* th, const float* x, const float* nodes, int64_t B,   \
                         int32_t k, int64_t Bg, float* loss, float* g, void* ws, size_t wsb, void* st) {           \
    return dgmk::Api<BK>::fredholm_step(d, th, x, nodes, B, k, Bg, loss, g, ws, wsb, st); }                        \
  int dgmk_jet_forward(const dgmk_net_desc* d, const float* th, const float* x, int64_t B, int32_t order,          \
                       float* Y, float* J, float* Hs, void* ws, size_t wsb, void* st) {                            \
    return dgmk::Api<BK>::jet_forward(d, th, x, B, order, Y, J, Hs, ws, wsb, st); }                                \
  int dgmk_jet_reverse(const dgmk_net_desc* d, const float* th, const float* x, int64_t B, int32_t order,          \
                       const float* gY, const float* gJ, const float* gHs, float* g, void* ws, size_t wsb,         \
                       void* st) {                                                                                 \
    return dgmk::Api<BK>::jet_reverse(d, th, x, B, order, gY, gJ, gHs, g, ws, wsb, st); }                          \
  int dgmk_eval(const dgmk_net_desc* d, const float* th, const float* x, int64_t B, float* Y, void* ws,            \
                size_t wsb, void* st) { return dgmk::Api<BK>::eval(d, th, x, B, Y, ws, wsb, st); }                 \
  int dgmk_adam(float* th, float* m, float* v, const float* g, const uint8_t* live, int64_t P, double lr,          \
                double b1, double b2, double eps, int64_t step, void* st) {                                        \
    return dgmk::Api<BK>::adam(th, m, v, g, live, P, lr, b1, b2, eps, step, st); }                                 \
  int dgmk_adam_dev(float* th, float* m, float* v, const float* g, const uint8_t* live, int64_t P, double lr,      \
                    double b1, double b2, double eps, long long* state, void* st) {                                \
    return dgmk::Api<BK>::adam_dev(th, m, v, g, live, P, lr, b1, b2, eps, state, st); }                            \
  int dgmk_sample_uniform(float* out, int64_t n, float lo, float hi, unsigned long long seed, uint32_t stream_id,  \
                          const long long* step_dev, long long step_add, void* st) {                               \
    return dgmk::Api<BK>::sample_uniform(out, n, lo, hi, seed, stream_id, step_dev, step_add, st); }               \
  int dgmk_sample_heat(float* X, float* X0, float* XBD1, float* XBD2, int64_t B, float xmax, float tmax,           \
                       float xbd2, unsigned long long seed, const long long* step_dev, long long step_add,         \
                       void* st) {                                                                                 \
    return dgmk::Api<BK>::sample_heat(X, X0, XBD1, XBD2, B, xmax, tmax, xbd2, seed, step_dev, step_add, st); }     \
  }
