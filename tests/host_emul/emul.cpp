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
struct HostBackend {
  bool bad_bt = false;
  int64_t hl_stride = 0;
  explicit HostBackend(void*) {}
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
