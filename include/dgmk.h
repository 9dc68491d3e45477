/* dgmk -- C ABI of the B200-native collocation training step.
 *
 * The reference (gdetor/differential_equations_dnn) is pure Python on torch: it
 * has no FFI layer.  The boundary this library replaces is the pair of Python
 * seams SURVEY.md 8(b) identifies; each entry point below cites the reference
 * code whose work it performs.  Callers are the torch.autograd.Functions in
 * differential_equations_dnn_b200/ (via ctypes); see INTEGRATION.md for the stub
 * a reference maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to FP32 data owned by the caller, except
 *    `desc` and where stated; the library allocates nothing and keeps no pointer
 *    after returning;
 *  - `theta` / `grad_theta` use the reference modules' named_parameters() order
 *    (dgmk_param_layout enumerates it);
 *  - work is enqueued on `stream` (a cudaStream_t passed as void*) and the call
 *    returns without synchronising; CUDA errors surface on the next call or on
 *    the caller's synchronise;
 *  - return value: 0 on success, negative DGMK_E* on failure with a message in
 *    dgmk_last_error() (thread-local).  Never exits, never throws;
 *  - re-entrant; concurrent calls on different streams/devices are safe as long
 *    as their workspaces are distinct.
 *  - There is no CPU implementation: host pointers are an error (DGMK_EDEVICE).
 */
#ifndef DGMK_H
#define DGMK_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGMK_VERSION 100

enum { DGMK_OK = 0, DGMK_EINVAL = -1, DGMK_EWORKSPACE = -2, DGMK_ECUDA = -3, DGMK_EDEVICE = -4 };

/* network families on the hot path */
enum {
  DGMK_MLP = 0,        /* neural_networks.MLP, no batch-norm (neural_networks.py:180-270) */
  DGMK_DGM_LINEAR = 1, /* dgm_net.DGM (dgm_net.py:71-119)                                 */
  DGMK_DGM_RAW = 2     /* neural_networks.DGM (neural_networks.py:130-177)                */
};
/* neural_networks.selectActivationFunction (neural_networks.py:24-41) */
enum { DGMK_RELU = 0, DGMK_SIGMOID = 1, DGMK_TANH = 2, DGMK_LEAKY_RELU = 3 };

/* workspace sizing classes */
enum {
  DGMK_WS_HEAT = 0, DGMK_WS_ODE = 1, DGMK_WS_FHN = 2, DGMK_WS_FREDHOLM = 3,
  DGMK_WS_JET0 = 4, DGMK_WS_JET1 = 5, DGMK_WS_JET2 = 6
};

typedef struct {
  int32_t kind;        /* DGMK_MLP / DGMK_DGM_LINEAR / DGMK_DGM_RAW          */
  int32_t input_dim;   /* d: 1 or 2                                          */
  int32_t output_dim;  /* o: 1..4                                            */
  int32_t hidden_size; /* H                                                  */
  int32_t num_layers;  /* L (same meaning as the reference constructors)     */
  int32_t activation;  /* MLP only; DGM_LINEAR is tanh, DGM_RAW is relu      */
  int32_t reserved[2];
} dgmk_net_desc;

int dgmk_version(void);
const char* dgmk_backend(void);    /* "cuda-sm100a" for the shipped library */
const char* dgmk_last_error(void);

/* Flat parameter layout = named_parameters() order of the reference module.
 * dgmk_param_count: total number of floats (includes neural_networks.DGM's dead
 * `dgm1` sub-layer, neural_networks.py:145).  dgmk_param_layout: tensor `index`
 * (0-based) -> offset/shape/live; returns the number of tensors when index < 0,
 * DGMK_EINVAL past the end. */
int64_t dgmk_param_count(const dgmk_net_desc* desc);
int dgmk_param_layout(const dgmk_net_desc* desc, int32_t index, int64_t* offset, int32_t* rows,
                      int32_t* cols, int32_t* live);

/* Recommended workspace size for `B` rows (Fredholm: k nodes per row).  The step
 * entry points process the batch in chunks sized to whatever workspace they are
 * given (any size >= dgmk_workspace_bytes(desc, cls, 1024, k) works); the jet
 * entry points need the full dgmk_workspace_bytes(desc, DGMK_WS_JETn, B, 0) because
 * the stash must survive from dgmk_jet_forward to dgmk_jet_reverse. */
size_t dgmk_workspace_bytes(const dgmk_net_desc* desc, int32_t ws_class, int64_t B, int32_t k);

/* ---- fused training steps: loss + d loss / d theta in one call ---------------
 * Each replaces `dgm_loss_func(...)` + `loss.backward()` of one reference script.
 * `B` rows are local to this call; every per-row term is scaled by 1/B_global so
 * that data-parallel ranks can SUM their (loss, grad).  `loss` is one
 * device float.  grad_theta of parameters the reference leaves at grad=None is 0. */

/* heat.py:50-95 -- x,x0,xbd1,xbd2: [B,2] (col0 = x, col1 = t); x_bd1,x_bd2: [B,1] */
int dgmk_heat_step(const dgmk_net_desc* desc, const float* theta, const float* x, const float* x0,
                   const float* xbd1, const float* xbd2, const float* x_bd1, const float* x_bd2,
                   int64_t B, int64_t B_global, float kappa, float* loss, float* grad_theta,
                   void* ws, size_t ws_bytes, void* stream);
/* simple_ode.py:41-63 with y = net(t), y0 = net(t0) (driver :98-101); t,t0,y_ic: [B,1] */
int dgmk_ode_step(const dgmk_net_desc* desc, const float* theta, const float* t, const float* t0,
                  const float* y_ic, int64_t B, int64_t B_global, float* loss, float* grad_theta,
                  void* ws, size_t ws_bytes, void* stream);
/* fitzhugh_nagumo.py:53-97; t,t0: [B,1], y_ic: [B,2]; net output_dim must be 2 */
int dgmk_fhn_step(const dgmk_net_desc* desc, const float* theta, const float* t, const float* t0,
                  const float* y_ic, int64_t B, int64_t B_global, float* loss, float* grad_theta,
                  void* ws, size_t ws_bytes, void* stream);
/* fredholm.py:47-74; x: [B,1]; nodes: [k,B,1], nodes[j] = the j-th `pi/2*rand_like(x)` */
int dgmk_fredholm_step(const dgmk_net_desc* desc, const float* theta, const float* x,
                       const float* nodes, int64_t B, int32_t k, int64_t B_global, float* loss,
                       float* grad_theta, void* ws, size_t ws_bytes, void* stream);

/* ---- module-level seam: net(x) with input derivatives --------------------------
 * Replaces `net(x)` followed by nested torch.autograd.grad(create_graph=True)
 * (heat.py:71-85).  order 0: Y only; 1: Y,J; 2: Y,J,Hs.  Y [B,o], J [B,o,d],
 * Hs [B,o,d,d] (symmetric).  dgmk_jet_reverse takes cotangents of all three (NULL =
 * zero) and ACCUMULATES nothing: grad_theta is overwritten. */
int dgmk_jet_forward(const dgmk_net_desc* desc, const float* theta, const float* x, int64_t B,
                     int32_t order, float* Y, float* J, float* Hs, void* ws, size_t ws_bytes,
                     void* stream);
int dgmk_jet_reverse(const dgmk_net_desc* desc, const float* theta, const float* x, int64_t B,
                     int32_t order, const float* gY, const float* gJ, const float* gHs,
                     float* grad_theta, void* ws, size_t ws_bytes, void* stream);
/* value-only forward in chunks (gridEvaluation, heat.py:152-172) */
int dgmk_eval(const dgmk_net_desc* desc, const float* theta, const float* x, int64_t B, float* Y,
              void* ws, size_t ws_bytes, void* stream);

/* ---- torch.optim.Adam(lr) defaults on the flat buffers (heat.py:115,141) --------
 * step = 1-based step count AFTER increment; live (uint8, may be NULL) = 0 for
 * parameters whose grad is None in the reference (skipped entirely).  The
 * hyper-parameters are doubles because torch forms 1-beta and the bias corrections
 * in Python doubles before touching FP32 tensors. */
int dgmk_adam(float* theta, float* m, float* v, const float* grad, const uint8_t* live, int64_t P,
              double lr, double beta1, double beta2, double eps, int64_t step, void* stream);

/* Same update with the step counter in device memory, for training loops captured in a CUDA graph
 * (nothing in the launch depends on host state).  state: 16 bytes of device memory, zero-initialised
 * by the caller = [int64 step count | 2 floats of scratch]; each call increments the count first. */
int dgmk_adam_dev(float* theta, float* m, float* v, const float* grad, const uint8_t* live, int64_t P,
                  double lr, double beta1, double beta2, double eps, long long* state, void* stream);

/* ---- on-device collocation sampler ---------------------------------------
 * Replaces the torch.rand / rand_like draws of the reference drivers (heat.py:125-126, simple_ode.py:91,
 * fitzhugh_nagumo.py:129, fredholm.py:67,100) with ONE launch per step of a counter-based generator, Philox4x32-10
 * (Salmon et al., SC'11): element i takes word i % 4 of the block with counter (lo32(i/4), hi32(i/4), lo32(step),
 * stream_id) under key (lo32(seed), hi32(seed)); u = (word >> 8) * 2^-24 in [0,1); value = lo + (hi - lo) * u
 * (two FP32 roundings).  step = *step_dev + step_add; step_dev (device int64, may be NULL) lets a captured CUDA graph
 * draw fresh points on every replay with no host work.  Statistical, not bitwise, parity with torch's stream; the
 * generator itself is pinned bit for bit by oracle/philox_np.py (Random123 known-answer vectors). */
int dgmk_sample_uniform(float* out, int64_t n, float lo, float hi, unsigned long long seed, uint32_t stream_id,
                        const long long* step_dev, long long step_add, void* stream);
/* heat.py:125-134 in one launch: x = xmax*u (stream 0), t = tmax*u' (stream 1); X = [x,t], X0 = [x,0],
 * XBD1 = [0,t], XBD2 = [xbd2,t], each [B,2] */
int dgmk_sample_heat(float* X, float* X0, float* XBD1, float* XBD2, int64_t B, float xmax, float tmax, float xbd2,
                     unsigned long long seed, const long long* step_dev, long long step_add, void* stream);

/* ---- diagnostics used by bench.py (not reference-facing) --------------------------- */
unsigned long long dgmk_launch_count(void); /* kernels launched by this library so far */
int dgmk_ffma_probe(const float* in, float* out, int blocks, int iters, void* stream);
/* 0: FP32 FFMA2 tiles only; 1 (default): tcgen05 3xTF32, fused GEMM + element-wise kernels and the
 * warp-specialised weight gradient where the shape allows; 2: tcgen05 3xTF32 streaming tiles +
 * separate element-wise kernels; 3: like 1 with the K = 3H data gradient of a DGM layer on the
 * streaming tile instead of the weight-resident kernel (dgmk_dgrad_res.cuh) */
void dgmk_set_gemm_engine(int engine);
/* 1 (default): hidden sizes <= 64 run the resident-tile step -- the whole step (forward jets, loss, reverse,
 * per-CTA gradient accumulation) in ONE persistent kernel with the activation stash and the packed weights in
 * shared memory (csrc/dgmk_tile.cuh); 0: the layer-wise path for every hidden size */
void dgmk_set_tile_engine(int on);
/* stage timeline of CTA 0 of the resident-tile kernel: (clock64, stage kind) pairs into buf (device, 2 * n int64) */
void dgmk_tile_profile(long long* buf, int n);
/* tiles per FP32 accumulation segment of the resident-tile step (default 256; <= 0 restores it) */
void dgmk_set_tile_flush(int tiles);
/* Per-kernel-class timing with CUDA events on the launch stream.  dgmk_profile(1) clears and
 * starts, dgmk_profile(0) stops; dgmk_profile_read synchronises the recorded events and returns the
 * class's summed duration [ms], launches, ALGORITHMIC flops and bytes.  Classes: 0 weight gradient,
 * 1 fused units-on-lanes GEMM + element-wise kernels, 2 streaming GEMM tiles, 3 element-wise, 4 reductions /
 * output layer, 5 resident-tile step kernel (hidden sizes <= 64). */
void dgmk_profile(int on);
int dgmk_profile_read(int cls, double* ms, long long* launches, double* flops, double* bytes);
/* Bt holds three [N,K] copies back to back: plain | tf32-hi | tf32-lo */
int dgmk_gemm_tc_probe(const float* A, const float* Bt, float* C, int64_t M, int N, int K, int64_t ld,
                       void* stream);
int dgmk_gemm_probe(const float* A, const float* B, float* C, int64_t M, int N, int K, int64_t ld,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DGMK_H */
