#!/usr/bin/env python
"""Benchmark of the collocation training step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # B200 arm (this repo's kernels)
  python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm: the reference's
        torch-autograd algorithm (oracle/ref_port.py -- the reference is pure Python and
        /root/reference does not exist on the GPU box) on all host cores, bounded sample

Workload (config.workload): BASELINE.json configs[1] -- heat.py's loss with dgm_net.DGM
(input_dim=2, output_dim=1, hidden_size=128, num_layers=3), 2^20 collocation rows per GPU
(one row = interior point + its IC and two BC companions), FP32, synthetic U[0,pi]x[0,3]
points, reference-seeded random-init weights.  A step = loss + d loss/d theta (fused
kernels) [+ all-reduce of the flat gradient at N>1] + fused Adam.  Weak scaling: the
per-GPU rows are fixed as N grows.

One JSON line on rank 0; see README "bench" for every key.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "collocation_rows_per_sec_training_step"
UNIT = "rows/s"
CPU_SAMPLE_ROWS = 4096


def f_alg(H, L, o=1, M=7, c=8):
    """Algorithmic FLOPs per row, SURVEY 8(d): 3 * M * (L*c*H^2 + 2*H*o)."""
    return 3 * M * (L * c * H * H + 2 * H * o)


# ------------------------------------------------------------------ clocks sampler
class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                    nv.nvmlDeviceGetCurrentClocksThrottleReasons
                bits = get(self.h)
                for b, name in self.REASONS.items():
                    if bits & b:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------ synthetic workload
def make_inputs(B, seed):
    """heat.py:125-134 on the CPU generator (identical bits for the CPU arm)."""
    gen = torch.Generator().manual_seed(seed)
    x = torch.pi * torch.rand([B, 1], generator=gen)
    t = 3.0 * torch.rand([B, 1], generator=gen)
    z = torch.zeros(B, 1)
    return (torch.cat([x, t], 1), torch.cat([x, z], 1), torch.cat([z, t], 1),
            torch.cat([z + torch.pi, t], 1), z.clone(), z.clone())


def cpu_reference_arm(H, L, steps, warmup, B_cpu=CPU_SAMPLE_ROWS):
    """The reference algorithm (nested torch.autograd.grad + backward + Adam) on the host."""
    from oracle import ref_port as rp
    from differential_equations_dnn_b200 import dgm_net
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1234)
    net = dgm_net.DGM(input_dim=2, output_dim=1, hidden_size=H, num_layers=L)
    theta = net.flat_theta().clone()
    spec = rp.NetSpec(rp.KIND_DGM_LINEAR, 2, 1, H, L, rp.ACT_TANH)
    inp = make_inputs(B_cpu, 1)
    m, v = torch.zeros_like(theta), torch.zeros_like(theta)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        loss, g = rp.loss_and_grad(rp.heat_loss, spec, theta, *inp)
        theta, m, v = rp.adam_step(theta, m, v, g, s + 1)
        float(loss)
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return {"value": B_cpu / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{B_cpu} rows/step of the same workload, {len(times)} timed steps after {warmup} warm-up, "
                      f"torch {torch.__version__} CPU autograd (oracle/ref_port.heat_loss + Adam), "
                      f"{cores} threads", "ms_per_step": dt * 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    H, L, B = a.hidden, a.layers, a.rows_per_gpu
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    config = {"workload": f"heat.py loss + dgm_net.DGM(2,1,{H},{L}), {B} rows/GPU "
                          f"(BASELINE configs[1]; row = interior + IC + 2 BC points)",
              "net": f"dgm_net.DGM(input_dim=2,output_dim=1,hidden_size={H},num_layers={L})",
              "rows_per_gpu": B, "global_rows": B * max(world, 1), "parallelism": f"dp{max(world, 1)}",
              "step": "fused loss+grad kernels, all-reduce(grad|loss) if N>1, fused Adam",
              "arithmetic": "FP32 in/out; GEMMs on tcgen05 with 3xTF32 split + RN chunk accumulation "
                            "(measured 1.2e-7 vs FP64, FP32 FFMA tile: 2.0e-7); element-wise jets in FP32",
              "l2": "per-step working set (activation stash, ~12 GB/chunk) >> 126 MB L2; inputs re-read from HBM"}

    if a.impl == "reference":
        if rank != 0:
            return
        W = max(a.warmup, 1)
        r = cpu_reference_arm(H, L, a.steps, W)
        config["workload"] += f"; CPU arm times a bounded sample of {CPU_SAMPLE_ROWS} rows per step"
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": W, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    import torch.distributed as dist
    import ctypes as C
    from differential_equations_dnn_b200 import _cabi, dgm_net, heat, optim, parallel, kernels

    lib = _cabi.load()  # raises if the CUDA library is missing: no fallback
    assert torch.cuda.is_available(), "bench.py (b200 arm) needs a GPU"
    # stdout carries exactly one JSON line: keep NCCL's own banner ("NCCL version ...", printed to stdout
    # when a box exports NCCL_DEBUG=VERSION) out of it
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    parallel.init_from_env("nccl")
    if dist.is_initialized():
        parallel.enable_data_parallel()
        world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", torch.cuda.current_device())
    W = max(a.warmup, 3)

    torch.manual_seed(1234)  # reference constructor order on the CPU generator, then move
    net = dgm_net.DGM(input_dim=2, output_dim=1, hidden_size=H, num_layers=L).to(dev)
    opt = optim.FusedAdam(net.parameters(), lr=1e-4)
    host = [t.pin_memory() for t in make_inputs(B, 1 + rank)]
    res = [t.to(dev) for t in host]
    stage = [torch.empty_like(t) for t in res]

    def barrier():
        if dist.is_initialized():
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        opt.zero_grad()
        loss = heat.dgm_loss_func(net, *res)
        loss.backward()
        opt.step()
        return loss

    def step_e2e():
        for d_, h_ in zip(stage, host):
            d_.copy_(h_, non_blocking=True)   # H2D of this step's rows from pinned memory
        opt.zero_grad()
        loss = heat.dgm_loss_func(net, *stage)
        loss.backward()
        opt.step()
        return loss.item()                    # D2H of the step's result

    def timed(fn, warm, steps, sample_clocks):
        for _ in range(warm):
            fn()
        barrier()
        sampler = ClockSampler(dev.index) if sample_clocks else None
        if sampler:
            sampler.start()
        l0 = lib.dgmk_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        launches = lib.dgmk_launch_count() - l0
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
        if dist.is_initialized():
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item(), t[1].item(), clocks, launches, out

    lib.dgmk_profile(1)   # CUDA events around every kernel class of the library, on the launch stream
    ms, wall_ms, clocks, launches, last = timed(step_resident, W, a.steps, True)
    lib.dgmk_profile(0)
    ms_step = ms / a.steps
    rows = B * max(world, 1)
    value = rows / (ms_step * 1e-3)
    prof = {}
    names = {0: "wg::wgrad_ws_kernel (weight gradient: warp-specialised tcgen05 kind::tf32, A^T in tensor memory, 3xTF32)",
             1: "lg::lane_gemm_kernel (fused GEMM + jet stage: weights in tensor memory, warp-specialised, 3xTF32)",
             2: "tc::gemm_nn_tc_kernel (streaming tcgen05 tile: K = 3H data gradient; FFMA2 tile for H % 128 != 0)",
             3: "ew_kernel<...> / rev1_e_kernel (stand-alone element-wise jet stages, loss, pack, Adam)",
             4: "wcolsum / rowdot / reduce_partials (output layer, column sums, second-stage reductions)"}
    for cls in range(5):
        t_, n_, f_, b_ = C.c_double(), C.c_longlong(), C.c_double(), C.c_double()
        _cabi.check(lib.dgmk_profile_read(cls, C.byref(t_), C.byref(n_), C.byref(f_), C.byref(b_)))
        per_step = (W + a.steps)   # the profile spans warm-up + timed steps: identical work per step
        if n_.value:
            prof[cls] = {"kernel": names[cls], "ms_per_step": t_.value / per_step, "launches_per_step": n_.value / per_step,
                         "alg_flops_per_step": f_.value / per_step, "alg_bytes_per_step": b_.value / per_step}
    e_ms, e_wall, _, _, _ = timed(step_e2e, 2, a.steps, False)
    e_step = max(e_ms, e_wall) / a.steps      # host-side copies/sync: take the larger clock
    e2e_value = rows / (e_step * 1e-3)

    if rank != 0:
        if dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel, from the timed region itself ------------------------
    # The library brackets every launch of a kernel class with CUDA events on its stream
    # (dgmk_profile) and sums the launch's ALGORITHMIC flops (2*M*N*K) and bytes (each operand /
    # result once).  The dominant class is the one with the largest summed duration.  Its tensor
    # roofline: peak = dense TF32 = 1/2 of the measured bf16 figure (sustained: timed inside a long
    # step); the kernels issue 3 TF32 MMAs per algorithmic product (FP32-grade accuracy), so their
    # own ceiling is peak/3 -- reported as frac_of_3xtf32_ceiling next to the HBM fraction.
    fl_row = f_alg(H, L)
    roof = None
    try:
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def timeit(fn, reps):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        # FP32 FFMA peak, measured live (MEASURED_PEAKS.json has no FP32 entry)
        pin = torch.ones(64, device=dev) * 1.0000001
        blocks, iters = 148 * 8, 20000
        pout = torch.empty(blocks * 256, device=dev)
        t_ms = timeit(lambda: lib.dgmk_ffma_probe(C.c_void_p(pin.data_ptr()), C.c_void_p(pout.data_ptr()), blocks, iters, st), 5)
        fp32_peak = 2.0 * 64 * iters * 256 * blocks / (t_ms * 1e-3) / 1e12
        peaks = {}
        mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(mp):
            peaks = json.load(open(mp))
        bf16 = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0))
        bf16_src = ("MEASURED_PEAKS.json bf16_tflops_sustained" if "bf16_tflops_sustained" in peaks else
                    "MEASURED_PEAKS.json bf16_tflops" if "bf16_tflops" in peaks else "fallback 1590 (B200_PROFILING.md)")
        hbm = peaks.get("hbm_gbs", 6500.0)
        tf32_peak = bf16 / 2.0
        kernels_ = {}
        for cls, p in prof.items():
            t = p["ms_per_step"] * 1e-3
            kernels_[str(cls)] = dict(p, share_of_step=p["ms_per_step"] / ms_step,
                                      achieved_tflops=p["alg_flops_per_step"] / t / 1e12 if p["alg_flops_per_step"] else None,
                                      achieved_gbs=p["alg_bytes_per_step"] / t / 1e9 if p["alg_bytes_per_step"] else None)
        gemm_cls = [c for c in prof if prof[c]["alg_flops_per_step"]]
        dom = max(gemm_cls, key=lambda c: prof[c]["ms_per_step"])
        k = kernels_[str(dom)]
        per_launch = 1.0 / k["launches_per_step"]
        tensor_frac3 = 3 * k["achieved_tflops"] / tf32_peak      # against the kernel's own 3xTF32 ceiling
        hbm_frac = k["achieved_gbs"] / hbm
        common = {"kernel": k["kernel"], "traffic": None,
                  "launch": {"avg_ms": k["ms_per_step"] * per_launch, "alg_flops": k["alg_flops_per_step"] * per_launch,
                             "alg_bytes": k["alg_bytes_per_step"] * per_launch, "per_step": k["launches_per_step"]},
                  "tensor_tflops_algorithmic": k["achieved_tflops"], "tensor_issue_tflops": 3 * k["achieved_tflops"],
                  "tf32_peak_tflops": tf32_peak, "frac_of_tf32_peak": k["achieved_tflops"] / tf32_peak,
                  "frac_of_3xtf32_ceiling": tensor_frac3,
                  "hbm_gbs": k["achieved_gbs"], "hbm_peak_gbs": hbm, "hbm_frac": hbm_frac,
                  "share_of_step": k["share_of_step"], "kernel_classes": kernels_}
        if hbm_frac >= tensor_frac3:   # the bound the kernel is closer to
            roof = dict(common, bound="hbm", achieved=k["achieved_gbs"], peak=hbm, unit="GB/s", frac=hbm_frac,
                        peak_source="MEASURED_PEAKS.json hbm_gbs; achieved = ALGORITHMIC bytes of the class's launches (every "
                                    "operand / result once: DESIGN.md section 3) / their CUDA-event durations inside the timed steps")
        else:
            roof = dict(common, bound="tensor", achieved=k["achieved_tflops"], peak=tf32_peak, unit="TFLOP/s",
                        frac=k["achieved_tflops"] / tf32_peak,
                        peak_source=f"TF32 dense = 1/2 x {bf16_src} = {tf32_peak:.1f} TFLOP/s; achieved = algorithmic 2MNK "
                                    "flops of the class's launches / their CUDA-event durations inside the timed steps "
                                    "(the kernel issues 3 TF32 MMAs per product: its own ceiling is peak/3)")
        # the north star's framing: whole step against the FP32 FFMA roofline of the reference path
        roof.update({"fp32_ffma_peak_live": fp32_peak, "alg_flops_per_row": fl_row,
                     "step_achieved": fl_row * B / (ms_step * 1e-3) / 1e12,
                     "step_frac_of_fp32_peak": fl_row * B / (ms_step * 1e-3) / 1e12 / fp32_peak})
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            tj = json.load(open(tr))
            key = {0: "wgrad_ws", 1: "lane_gemm", 2: "gemm_nn_tc"}.get(dom)
            ratio = tj.get(key + "_dram_over_algorithmic")
            if ratio is not None:   # ncu dram__bytes_read+write per launch / algorithmic bytes of that launch
                roof["traffic"] = ratio * roof["launch"]["alg_bytes"]
                roof["traffic_source"] = tj.get(key + "_source")
    except Exception as e:  # the number above is still valid without the probe
        roof = {"error": repr(e)}

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        r = cpu_reference_arm(H, L, 3, 1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    h2d = sum(t.numel() * 4 for t in host)
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": max(world, 1), "steps": a.steps, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config, "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": e_step},
        "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
        "loss_last": float(last.detach()) if hasattr(last, "detach") else float(last), "wall_ms_per_step": wall_ms / a.steps}), flush=True)
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
