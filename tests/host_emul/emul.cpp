// TEST INFRASTRUCTURE ONLY -- never shipped, never loaded by the package.
//
// Instantiates the SAME orchestration templates as the CUDA library
// (csrc/dgmk_pipeline.h, dgmk_capi_impl.h, dgmk_ops.h) with a backend made of plain
// host loops, so that tests can check buffer carving, layer ordering, packing and
// scaling against the oracle in the GPU-less build container.  The GEMM tiles and
// reduction kernels themselves are CUDA-only and are covered by the -m gpu tests.
// Built by tests/host_emul/build.sh into tests/host_emul/libdgmk_emul.so.
#include <vector>
#include "../../differential_equations_dnn_b200/csrc/dgmk_capi_impl.h"

namespace {
int g_fred_nodes = 0, g_fred_points = 0;   // dgmk_emul_set_fredholm_blocks: the sub-tiled Fredholm body (dgmk_steps.h)
bool g_inplace_rev = false;   // dgmk_emul_set_inplace: run the reverse pass in the resident-tile step's in-place mode
struct HostBackend : dgmk::BackendTraitsAll {
  bool inplace_rev() const { return g_inplace_rev; }
  int fredholm_block_nodes() const { return g_fred_nodes; }
  int fredholm_block_points() const { return g_fred_points; }
  bool bad_bt = false;
  int64_t hl_stride = 0;
  explicit HostBackend(void*) {}
  void note_bytes(double) {}
  template <class F> void ew(const F& f, int64_t n) { for (int64_t i = 0; i < n; ++i) f(i); }
  // the four-units-per-thread form of the same functors: exercise it on the host too
  template <class F> void ew4(const F& f, int64_t n) { for (int64_t k = 0; k < n / 4; ++k) f.vec4(k); }
  // the fused GEMM + element-wise kernels are CUDA-only: the harness always takes the unfused route
  bool lane_ok(int, int) const { return false; }
  template <class CS, int ACT>
  void dgm_fwd_fused(const dgmk::XSrc&, const float*, float*, const dgmk::F4*, float*, float*, const float*, int, int64_t) {}
  template <class CS, int ACT>
  void dgm_rev2_fused(const float*, const float*, float*, float*, const float*, int, int64_t) {}
  template <class CS, class F>
  void dgm_rev1_e(const F&, const dgmk::XSrc&, int64_t, float*, float*, int64_t) {}
  template <class CS, int ACT>
  void mlp_fwd_fused(const float*, float*, const dgmk::F4*, float*, const float*, int, int64_t) {}
  void lane_store(const float*, int64_t, const float*, float*, int64_t, int, int64_t) {}
  template <class CS, int ACT>
  void mlp_rev_fused(const float*, const float*, const float*, float*, int, int64_t) {}
  template <class CS, int ACT>
  void input_rev_fused(const float*, const float*, const dgmk::F4*, const float*, float*, int, int64_t) {}
  void gemm_nn(const float* A, int64_t lda, const float* B, int64_t ldb, const float* Bt, int64_t ldbt, float* C, int64_t ldc, int64_t M,
               int N, int K, bool acc) {
    // both packed orientations must describe the same matrix (checks the packing tables)
    for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) {
      float b = B[(int64_t)k * ldb + n]; const float* t = Bt + (int64_t)n * ldbt + k;
      if (b != t[0] || t[hl_stride] + t[2 * hl_stride] != b || t[hl_stride] != dgmk::tf32_round(b)) bad_bt = true;
    }
    for (int64_t m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        float s = 0.f;
        for (int k = 0; k < K; ++k) s = fmaf(A[m * lda + k], B[(int64_t)k * ldb + n], s);
        C[m * ldc + n] = acc ? C[m * ldc + n] + s : s;
      }
  }
  void gemm_tn_acc(const float* A, int64_t lda, const float* S, int64_t lds, float* out, int N, int Kd, int64_t M,
                   const float* E, float* outE, int64_t ldoE, float*, int64_t) {
    std::vector<double> acc((size_t)N * Kd, 0.0), eacc((size_t)4 * N, 0.0);
    for (int64_t m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double a = A[m * lda + n];
        for (int k = 0; k < Kd; ++k) acc[(size_t)n * Kd + k] += a * S[m * lds + k];
        if (E) for (int e = 0; e < 4; ++e) eacc[(size_t)e * N + n] += a * E[m * 4 + e];
      }
    for (size_t i = 0; i < acc.size(); ++i) out[i] += (float)acc[i];
    if (E) for (int e = 0; e < 4; ++e) for (int n = 0; n < N; ++n) outE[e * ldoE + n] += (float)eacc[(size_t)e * N + n];
  }
  void wcolsum_acc(const float* Mat, int64_t ldm, int N, const float* Wt, int64_t M, float* out, float*, int64_t, int64_t ldo = 0) {
    int NE = Wt ? 4 : 1;
    std::vector<double> acc((size_t)NE * N, 0.0);
    for (int64_t r = 0; r < M; ++r)
      for (int e = 0; e < NE; ++e) {
        double w = Wt ? Wt[r * 4 + e] : 1.0;
        for (int n = 0; n < N; ++n) acc[(size_t)e * N + n] += w * Mat[r * ldm + n];
      }
    for (int e = 0; e < NE; ++e) for (int n = 0; n < N; ++n) out[(ldo > 0 ? e * ldo : (int64_t)e * N) + n] += (float)acc[(size_t)e * N + n];
  }
  void rowdot(const float* S, int64_t lds, const float* W, const float* b, float* U, int64_t M, int Hp, int o, int C) {
    for (int64_t r = 0; r < M; ++r)
      for (int m = 0; m < 4; ++m) {
        float s = 0.f;
        if (m < o) {
          for (int j = 0; j < Hp; ++j) s += S[r * lds + j] * W[m * Hp + j];
          if (r % C == 0) s += b[m];
        }
        U[r * 4 + m] = s;
      }
  }
  void zero(void* p, size_t bytes) { memset(p, 0, bytes); }
  void copy(void* dst, const void* src, size_t bytes) { memcpy(dst, src, bytes); }
  bool is_device_ptr(const void*) { return true; }
  const char* error() { return bad_bt ? "Bt != B^T" : nullptr; }
};
}  // namespace

DGMK_DEFINE_C_API(HostBackend, "host-emulation (tests only)")
extern "C" void dgmk_emul_set_inplace(int on) { g_inplace_rev = on != 0; }
extern "C" void dgmk_emul_set_fredholm_blocks(int points, int nodes) { g_fred_points = points; g_fred_nodes = nodes; }

// ---- unit checks of functor code that only the CUDA backend's fused path calls (same templates, host
// instantiation): the structural input-map adjoint and the V-units-per-thread DgmRev1 stage ------------------
namespace {
struct Lcg {   // deterministic, platform-independent
  uint64_t s;
  explicit Lcg(unsigned seed) : s(seed * 2654435761u + 12345u) {}
  float uni() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (float)((s >> 40) & 0xFFFFFF) / 16777216.0f; }
};

// max |input_map_adj sums - Abar^T E| / (1 + max |Abar^T E|) over random points (E rows from ExtInputFn)
template <class CS>
double check_input_map_adj(int npoints, int d, unsigned seed) {
  Lcg r(seed);
  std::vector<float> X((size_t)npoints * d), E((size_t)npoints * CS::C * 4);
  for (auto& v : X) v = 3.0f * r.uni() - 1.0f;
  dgmk::XSrc xs = dgmk::xsrc1(X.data(), npoints, d);
  dgmk::ExtInputFn<CS> fe; fe.xs = xs; fe.E = E.data();
  for (int p = 0; p < npoints; ++p) fe(p);
  float g[3] = {0.f, 0.f, 0.f};
  double ref[4] = {0, 0, 0, 0};
  for (int p = 0; p < npoints; ++p) {
    float ab[CS::C];
    for (int c = 0; c < CS::C; ++c) ab[c] = 2.0f * r.uni() - 1.0f;
    const float* x = xs.at(p);
    dgmk::input_map_adj<CS>(ab, x[0], d > 1 ? x[1] : 0.f, g);
    for (int c = 0; c < CS::C; ++c)
      for (int e = 0; e < 4; ++e) ref[e] += (double)ab[c] * E[((size_t)p * CS::C + c) * 4 + e];
  }
  double worst = std::fabs(ref[3]), scale = 1.0;   // the fourth E column is structurally zero
  for (int e = 0; e < 3; ++e) scale = std::fmax(scale, 1.0 + std::fabs(ref[e]));
  for (int e = 0; e < 3; ++e) worst = std::fmax(worst, std::fabs((double)g[e] - ref[e]));
  return worst / scale;
}

template <class CS, int V>
struct SumSink {
  float x0, x1; double (*g)[3][3];
  void operator()(int u, int slot, const float* ab) const {
    float t[3] = {0.f, 0.f, 0.f};
    dgmk::input_map_adj<CS>(ab, x0, x1, t);
    for (int e = 0; e < 3; ++e) g[u][slot == 3 ? 2 : slot][e] += t[e];
  }
};
// DgmRev1Fn::runv<V> against operator(): 0 = stored cotangents bit-identical and the sink saw exactly them,
// 1 = stored values differ, 2 = sink sums differ
template <class CS, int ACT, int V>
int check_rev1_runv(int npoints, unsigned seed) {
  constexpr int Hp = 8, C = CS::C;
  Lcg r(seed);
  const size_t M = (size_t)npoints * C;
  std::vector<float> A4(M * 4 * Hp), S(M * Hp), SBn(M * Hp), ABa(M * 4 * Hp, -7.f), ABb(M * 4 * Hp, -7.f), SPa(M * Hp), SPb(M * Hp);
  for (size_t i = 0; i < A4.size(); ++i) {
    const bool value_row = ((i / (4 * Hp)) % C) == 0;
    A4[i] = value_row ? (ACT == dgmk::ACT_TANH ? 1.8f * r.uni() - 0.9f : (r.uni() < 0.3f ? 0.f : r.uni())) : 2.0f * r.uni() - 1.0f;
  }
  for (auto& v : S) v = 2.0f * r.uni() - 1.0f;
  for (auto& v : SBn) v = 2.0f * r.uni() - 1.0f;
  dgmk::DgmRev1Fn<CS, ACT> fa; fa.A4 = A4.data(); fa.S = S.data(); fa.SBn = SBn.data(); fa.AB4 = ABa.data(); fa.SBp = SPa.data(); fa.Hp = Hp;
  dgmk::DgmRev1Fn<CS, ACT> fb = fa; fb.AB4 = ABb.data(); fb.SBp = SPb.data();
  for (int64_t i = 0; i < (int64_t)npoints * Hp; ++i) fa(i);
  double g[V][3][3] = {};
  for (int64_t k = 0; k < (int64_t)npoints * (Hp / V); ++k) {
    SumSink<CS, V> sink; sink.x0 = 0.25f * (float)(k % 7); sink.x1 = -0.5f; sink.g = g;
    fb.template runv<V>(k, sink);
  }
  if (memcmp(ABa.data(), ABb.data(), ABa.size() * 4) || memcmp(SPa.data(), SPb.data(), SPa.size() * 4)) return 1;
  // what the sink was shown must be what was stored: redo its sums from the stored cotangents
  double h[V][3][3] = {};
  for (int64_t k = 0; k < (int64_t)npoints * (Hp / V); ++k) {
    const int64_t p = k / (Hp / V); const int j0 = (int)(k % (Hp / V)) * V;
    for (int u = 0; u < V; ++u)
      for (int slot : {0, 1, 3}) {
        float ab[C], t[3] = {0.f, 0.f, 0.f};
        for (int c = 0; c < C; ++c) ab[c] = ABa[((size_t)p * C + c) * 4 * Hp + slot * Hp + j0 + u];
        dgmk::input_map_adj<CS>(ab, 0.25f * (float)(k % 7), -0.5f, t);
        for (int e = 0; e < 3; ++e) h[u][slot == 3 ? 2 : slot][e] += t[e];
      }
  }
  for (int u = 0; u < V; ++u) for (int a = 0; a < 3; ++a) for (int e = 0; e < 3; ++e)
    if (g[u][a][e] != h[u][a][e]) return 2;
  return 0;
}
}  // namespace

extern "C" {
double dgmk_emul_check_input_map_adj(int cs, int npoints, int d, unsigned seed) {
  switch (cs) {
    case dgmk::CS_V: return check_input_map_adj<dgmk::CsV>(npoints, d, seed);
    case dgmk::CS_D1O1: return check_input_map_adj<dgmk::CsD1O1>(npoints, d, seed);
    case dgmk::CS_HEAT: return check_input_map_adj<dgmk::CsHeat>(npoints, d, seed);
    case dgmk::CS_D2O1: return check_input_map_adj<dgmk::CsD2O1>(npoints, d, seed);
    case dgmk::CS_D1O2: return check_input_map_adj<dgmk::CsD1O2>(npoints, d, seed);
    default: return check_input_map_adj<dgmk::CsD2O2>(npoints, d, seed);
  }
}
int dgmk_emul_check_rev1_runv(int cs, int tanh_gates, int V, int npoints, unsigned seed) {
#define DGMK_RV(CS)                                                                                         \
  return tanh_gates ? (V == 4 ? check_rev1_runv<CS, dgmk::ACT_TANH, 4>(npoints, seed) : check_rev1_runv<CS, dgmk::ACT_TANH, 2>(npoints, seed)) \
                    : (V == 4 ? check_rev1_runv<CS, dgmk::ACT_RELU, 4>(npoints, seed) : check_rev1_runv<CS, dgmk::ACT_RELU, 2>(npoints, seed));
  switch (cs) {
    case dgmk::CS_V: DGMK_RV(dgmk::CsV)
    case dgmk::CS_D1O1: DGMK_RV(dgmk::CsD1O1)
    default: DGMK_RV(dgmk::CsHeat)
  }
#undef DGMK_RV
}
}
