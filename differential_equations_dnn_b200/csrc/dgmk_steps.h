// One CHUNK of a fused training step: forward of the chunk's passes, loss + cotangent seeds, reverse.
//
// The same bodies run in two places:
//   * on the host, inside the chunk loops of dgmk_capi_impl.h (each stage = one kernel launch over a chunk that is
//     sized to the caller's workspace in HBM), and
//   * on the device, inside the persistent tile kernels of dgmk_tile.cuh (each stage = one CTA-cooperative call
//     over a chunk -- a TILE of points -- that lives in shared memory).
// Reference: heat.py:71-95, simple_ode.py:54-63, fitzhugh_nagumo.py:69-97 (+ loss.backward()).
#pragma once
#include "dgmk_pipeline.h"

namespace dgmk {

struct HeatArgs {
  const float* x; const float* x0; const float* xbd1; const float* xbd2; const float* t_bd1; const float* t_bd2;
  float kappa, inv;
};
// rows [p0, p0 + r) of the batch; cv.off is reset to `mark` for each pass (the passes reuse the chunk region)
DGMK_NOCHECK
template <class PIPE>
DGMK_HD_PLAIN bool heat_chunk(PIPE& P, Carver& cv, size_t mark, const HeatArgs& a, int64_t p0, int64_t r) {
  auto& c = P.c;
  auto& bk = P.bk;
  {  // interior rows: u_t - kappa u_xx  (heat.py:71-87)
    cv.off = mark;
    PassBufs pb; RevBufs rb;
    pb.xs = xsrc1(a.x + p0 * 2, r, 2);
    bk.stage_coords(pb.xs, r);
    if (!carve_pass(cv, c.n, &pb, r, CS_HEAT) || !carve_rev(cv, c.n, &rb, pb.M, bk.inplace_rev())) return false;
    P.forward(pb);
    HeatInteriorFn f; f.U = pb.U; f.UB = pb.UB; f.Lp = c.Lp; f.kappa = a.kappa; f.inv = a.inv;
    bk.ew(f, r);
    P.reverse(pb, rb, r);
  }
  {  // companions: IC row (x,0), BC rows (0,t) and (pi,t)  (heat.py:89-94)
    cv.off = mark;
    PassBufs pb; RevBufs rb;
    pb.xs.p[0] = a.x0 + p0 * 2; pb.xs.p[1] = a.xbd1 + p0 * 2; pb.xs.p[2] = a.xbd2 + p0 * 2;
    pb.xs.block_rows = r; pb.xs.block_stride = 0; pb.xs.nptr = 3; pb.xs.d = 2;
    bk.stage_coords(pb.xs, 3 * r);
    if (!carve_pass(cv, c.n, &pb, 3 * r, CS_V) || !carve_rev(cv, c.n, &rb, pb.M, bk.inplace_rev())) return false;
    P.forward(pb);
    ValueTargetFn f; f.U = pb.U; f.UB = pb.UB; f.Lp = c.Lp; f.xs = pb.xs;
    f.tgt[0] = nullptr; f.tgt[1] = a.t_bd1 + p0; f.tgt[2] = a.t_bd2 + p0;
    f.mode[0] = 1; f.mode[1] = 0; f.mode[2] = 0; f.o = 1; f.inv = a.inv;
    bk.ew(f, 3 * r);
    P.reverse(pb, rb, 3 * r);
  }
  return true;
}

struct OdeArgs {
  const float* t; const float* t0; const float* y_ic;
  float inv; int32_t fhn;
};
DGMK_NOCHECK
template <class PIPE>
DGMK_HD_PLAIN bool ode_like_chunk(PIPE& P, Carver& cv, size_t mark, const OdeArgs& a, int64_t p0, int64_t r) {
  auto& c = P.c;
  auto& bk = P.bk;
  {
    cv.off = mark;
    PassBufs pb; RevBufs rb;
    pb.xs = xsrc1(a.t + p0, r, 1);
    bk.stage_coords(pb.xs, r);
    if (!carve_pass(cv, c.n, &pb, r, CS_D1O1) || !carve_rev(cv, c.n, &rb, pb.M, bk.inplace_rev())) return false;
    P.forward(pb);
    if (a.fhn) {
      FhnInteriorFn f; f.U = pb.U; f.UB = pb.UB; f.Lp = c.Lp; f.I = 0.5f; f.alpha = 0.7f; f.beta = 0.8f; f.tau = 2.5f; f.inv = a.inv;
      bk.ew(f, r);
    } else {
      OdeInteriorFn f; f.U = pb.U; f.UB = pb.UB; f.Lp = c.Lp; f.inv = a.inv;
      bk.ew(f, r);
    }
    P.reverse(pb, rb, r);
  }
  {  // initial-condition rows (simple_ode.py:62; fitzhugh_nagumo.py:95 -- mean over 2B elements)
    cv.off = mark;
    PassBufs pb; RevBufs rb;
    pb.xs = xsrc1(a.t0 + p0, r, 1);
    bk.stage_coords(pb.xs, r);
    if (!carve_pass(cv, c.n, &pb, r, CS_V) || !carve_rev(cv, c.n, &rb, pb.M, bk.inplace_rev())) return false;
    P.forward(pb);
    ValueTargetFn f; f.U = pb.U; f.UB = pb.UB; f.Lp = c.Lp; f.xs = pb.xs;
    f.tgt[0] = a.y_ic + p0 * c.n.o; f.tgt[1] = f.tgt[2] = nullptr; f.mode[0] = f.mode[1] = f.mode[2] = 0;
    f.o = c.n.o; f.inv = a.fhn ? a.inv * 0.5f : a.inv;
    bk.ew(f, r);
    P.reverse(pb, rb, r);
  }
  return true;
}

struct FredArgs {
  const float* x; const float* nodes;   // x [B], nodes [k][B]
  int64_t B; int32_t k; float dr, inv;
};
// One block of r points [p0, p0 + r) with ALL their k Monte-Carlo nodes (fredholm.py:64-74), the node rows walked
// in sub-tiles of J nodes (r * J rows) so that the stash fits a small memory (shared memory in the resident-tile
// step).  The integral needs every node value before any node row can be seeded, so the node rows are evaluated
// twice when they do not fit one sub-tile: once forward-only (running integral), once forward + reverse.
// Ip: r floats of scratch (running integral, then g = 2 r / B).  Same arithmetic and summation order as FredholmFn.
DGMK_NOCHECK
template <class PIPE>
DGMK_HD_PLAIN bool fredholm_block(PIPE& P, Carver& cv, size_t mark, const FredArgs& a, int64_t p0, int64_t r, int32_t J, float* Ip) {
  auto& c = P.c;
  auto& bk = P.bk;
  cv.off = mark;
  PassBufs px, pn; RevBufs rb;
  px.xs = xsrc1(a.x + p0, r, 1);
  if (!carve_pass(cv, c.n, &px, r, CS_V)) return false;
  const size_t mark_n = cv.off;
  P.forward(px);
  bk.zero(Ip, (size_t)r * 4);
  const int nsub = (a.k + J - 1) / J;
  for (int32_t j0 = 0; j0 < a.k; j0 += J) {   // pass 1 over the nodes: values only
    const int32_t jj = (a.k - j0 < J) ? a.k - j0 : J;
    cv.off = mark_n;
    pn.xs = xsrc1(a.nodes + p0 + (int64_t)j0 * a.B, r, 1);
    pn.xs.block_stride = a.B;
    if (!carve_pass(cv, c.n, &pn, r * jj, CS_V) || !carve_rev(cv, c.n, &rb, pn.M, bk.inplace_rev())) return false;
    bk.stage_coords(pn.xs, r * jj);
    P.forward(pn);
    FredAccFn f; f.Un = pn.U; f.x = a.x + p0; f.T = a.nodes + p0 + (int64_t)j0 * a.B; f.Ip = Ip; f.rows = r; f.Tstride = a.B; f.jj = jj;
    bk.ew(f, r);
  }
  {
    FredResFn f; f.Ux = px.U; f.UBx = px.UB; f.Lp = c.Lp; f.x = a.x + p0; f.Ip = Ip; f.dr = a.dr; f.inv = a.inv;
    bk.ew(f, r);
    P.add_loss(r);
  }
  for (int32_t j0 = 0; j0 < a.k; j0 += J) {   // pass 2: seeds + reverse (forward again unless the one sub-tile is still resident)
    const int32_t jj = (a.k - j0 < J) ? a.k - j0 : J;
    cv.off = mark_n;
    pn.xs = xsrc1(a.nodes + p0 + (int64_t)j0 * a.B, r, 1);
    pn.xs.block_stride = a.B;
    if (!carve_pass(cv, c.n, &pn, r * jj, CS_V) || !carve_rev(cv, c.n, &rb, pn.M, bk.inplace_rev())) return false;
    if (nsub > 1) { bk.stage_coords(pn.xs, r * jj); P.forward(pn); }
    FredSeedFn f; f.UBn = pn.UB; f.x = a.x + p0; f.T = a.nodes + p0 + (int64_t)j0 * a.B; f.G = Ip; f.rows = r; f.Tstride = a.B; f.dr = a.dr;
    bk.ew(f, r * jj);
    P.reverse(pn, rb);
  }
  // the x rows last: their stash sits in front of the node region and is untouched by it
  cv.off = mark_n;
  if (!carve_rev(cv, c.n, &rb, px.M, bk.inplace_rev())) return false;
  P.reverse(px, rb);
  return true;
}

}  // namespace dgmk
