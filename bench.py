#!/usr/bin/env python
"""Benchmark of the collocation training step (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W                    # B200 arm (this repo's kernels)
  python bench.py --impl reference --gpus N --steps K --warmup W   # reference arm: the UNMODIFIED reference
        (oracle/_ref, placed there by oracle/vendor_ref.py; oracle/ref_port.py if that is absent) on the
        box's host cores, bounded sample of the same workload, no GPU, nothing of this package imported

  --config heat (default) | ode | fhn | fredholm   one line per BASELINE config
  --rows-per-gpu R        rows per GPU (weak scaling: fixed as N grows); default per config
  --scaling strong --global-rows G   fixed global batch split over the ranks (config 5: G = 2^24)

Default workload (config.workload): BASELINE.json configs[1] -- heat.py's loss with dgm_net.DGM(input_dim=2,
output_dim=1, hidden_size=128, num_layers=3), 2^20 collocation rows per GPU (one row = interior point + its IC
and two BC companions), FP32, synthetic U[0,pi]x[0,3] points, reference-seeded random-init weights.  A step =
loss + d loss/d theta (fused kernels) [+ all-reduce of the flat gradient at N>1] + fused Adam.

One JSON line on rank 0; see README "bench" for every key.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import math
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="heat", choices=["heat", "ode", "fhn", "fredholm"])
    ap.add_argument("--net", default=None, help="fhn: mlp (default, the BASELINE config) | dgm (as shipped); heat: dgm | mlp")
    ap.add_argument("--rows-per-gpu", type=int, default=None)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--global-rows", type=int, default=None)
    ap.add_argument("--hidden", type=int, default=None)
    ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--k", type=int, default=1024, help="fredholm: quadrature nodes per point")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cuda-eager", action="store_true")
    ap.add_argument("--no-driver-latency", action="store_true")
    return ap.parse_args()


ARGS = parse_args()
if ARGS.impl == "reference":
    os.environ["CUDA_VISIBLE_DEVICES"] = ""   # the reference arm is the CPU implementation: no GPU in this process

import torch  # noqa: E402

METRIC = "collocation_rows_per_sec_training_step"
UNIT = "rows/s"


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


@contextlib.contextmanager
def quiet():
    """The reference constructors print ("No batch normalization"): keep stdout to the one JSON line."""
    with contextlib.redirect_stdout(sys.stderr):
        yield


# ------------------------------------------------------------------ workloads
class Workload:
    """One BASELINE config: synthetic inputs exactly as the reference driver builds them (CPU generator, so the
    reference arm and the B200 arm see identical bits), the network, and the step through either API."""
    name = ""
    M = 1            # value-equivalent forward passes per row
    in_bytes = 0     # algorithmic HBM bytes per row
    default_rows = 1 << 20

    def f_alg(self):
        """Algorithmic FLOPs per row, SURVEY 8(d): 3 * M * (L*c*H^2 + 2*H*o)."""
        c = 2 if self.kind == "mlp" else 8
        return 3 * self.M * (self.L * c * self.H * self.H + 2 * self.H * self.o)

    def net_name(self):
        if self.kind == "mlp":
            return f"neural_networks.MLP(input_dim={self.d},output_dim={self.o},hidden_size={self.H},num_layers={self.L},activation='{self.act}')"
        if self.kind == "dgm":
            return f"dgm_net.DGM(input_dim={self.d},output_dim={self.o},hidden_size={self.H},num_layers={self.L})"
        return f"neural_networks.DGM(input_dim={self.d},output_dim={self.o},hidden_size={self.H},num_layers={self.L})"

    def build_net(self, mods):
        """mods: namespace with .neural_networks / .dgm_net (the product package or the reference)."""
        torch.manual_seed(1234)   # reference constructor order on the CPU generator
        with quiet():
            if self.kind == "mlp":
                return mods.neural_networks.MLP(input_dim=self.d, output_dim=self.o, hidden_size=self.H,
                                                num_layers=self.L, activation=self.act)
            if self.kind == "dgm":
                return mods.dgm_net.DGM(input_dim=self.d, output_dim=self.o, hidden_size=self.H, num_layers=self.L)
            return mods.neural_networks.DGM(input_dim=self.d, output_dim=self.o, hidden_size=self.H, num_layers=self.L)


class Heat(Workload):
    name, file, M, in_bytes = "heat", "heat.py", 7, 40

    def __init__(self, a):
        self.kind = a.net or "dgm"
        self.d, self.o, self.H, self.L, self.act = 2, 1, a.hidden or 128, a.layers or 3, "tanh"
        self.tag = "BASELINE configs[1]; row = interior + IC + 2 BC points"

    def make_inputs(self, B, seed):
        """heat.py:125-134"""
        gen = torch.Generator().manual_seed(seed)
        x = torch.pi * torch.rand([B, 1], generator=gen)
        t = 3.0 * torch.rand([B, 1], generator=gen)
        z = torch.zeros(B, 1)
        return [torch.cat([x, t], 1), torch.cat([x, z], 1), torch.cat([z, t], 1),
                torch.cat([z + torch.pi, t], 1), z.clone(), z.clone()]

    def product_loss(self, pk, net, inp):
        return pk.heat.dgm_loss_func(net, *inp)

    def reference_loss(self, ref, net, inp):
        inp[0].requires_grad_(True)          # heat.py:129
        return ref.heat.dgm_loss_func(net, *inp)

    def port_loss(self, rp):
        return rp.heat_loss


class Ode(Workload):
    name, file, M, in_bytes = "ode", "simple_ode.py", 3, 12

    def __init__(self, a):
        self.kind = a.net or "mlp"
        self.d, self.o, self.H, self.L, self.act = 1, 1, a.hidden or 32, a.layers or 1, "relu"
        self.tag = "BASELINE configs[0] at the throughput size; row = interior + IC point"

    def make_inputs(self, B, seed):
        """simple_ode.py:87-94"""
        gen = torch.Generator().manual_seed(seed)
        return [1.01 * torch.rand([B, 1], generator=gen), torch.zeros(B, 1), 2.0 * torch.ones(B, 1)]

    def product_loss(self, pk, net, inp):
        with pk.simple_ode.deferred_forward(net):
            y, y0 = net(inp[0]), net(inp[1])
        return pk.simple_ode.dgm_loss_func(y, y0, inp[0], inp[2])

    def reference_loss(self, ref, net, inp):
        inp[0].requires_grad_(True)          # simple_ode.py:93
        return ref.simple_ode.dgm_loss_func(net(inp[0]), net(inp[1]), inp[0], inp[2])

    def port_loss(self, rp):
        return rp.ode_loss


class Fhn(Ode):
    name, file, M, in_bytes = "fhn", "fitzhugh_nagumo.py", 3, 16

    def __init__(self, a):
        self.kind = a.net or "mlp"
        if self.kind == "mlp":
            self.d, self.o, self.H, self.L, self.act = 1, 2, a.hidden or 128, a.layers or 3, "tanh"
        else:
            self.d, self.o, self.H, self.L, self.act = 1, 2, a.hidden or 128, a.layers or 4, "tanh"
        self.tag = "BASELINE configs[2]; row = interior + IC point"

    def make_inputs(self, B, seed):
        """fitzhugh_nagumo.py:121,129 (the sampler that scales past 200 rows), :218"""
        gen = torch.Generator().manual_seed(seed)
        return [30.01 * torch.rand([B, 1], generator=gen), torch.zeros(B, 1), torch.zeros(B, 2)]

    def product_loss(self, pk, net, inp):
        with pk.fitzhugh_nagumo.deferred_forward(net):
            y, y0 = net(inp[0]), net(inp[1])
        return pk.fitzhugh_nagumo.dgm_loss_func(y, y0, inp[0], inp[2])

    def reference_loss(self, ref, net, inp):
        inp[0].requires_grad_(True)
        return ref.fitzhugh_nagumo.dgm_loss_func(net(inp[0]), net(inp[1]), inp[0], inp[2])

    def port_loss(self, rp):
        return rp.fhn_loss


class Fredholm(Workload):
    name, file = "fredholm", "fredholm.py"
    default_rows = 1 << 14

    def __init__(self, a):
        self.kind = a.net or "dgmraw"
        self.d, self.o, self.H, self.L, self.act = 1, 1, a.hidden or 32, a.layers or 1, "relu"
        self.k = a.k
        self.M, self.in_bytes = self.k + 1, 4 * (self.k + 1)
        self.tag = f"BASELINE configs[3]; row = point + its k = {self.k} Monte-Carlo quadrature nodes"

    def make_inputs(self, B, seed):
        """fredholm.py:100 and the k rand_like draws of :66-67, in loop order"""
        gen = torch.Generator().manual_seed(seed)
        x = (math.pi / 2) * torch.rand([B, 1], generator=gen)
        return [x, (math.pi / 2) * torch.rand([self.k, B, 1], generator=gen)]

    def product_loss(self, pk, net, inp):
        return pk.fredholm.dgm_loss_func(net, inp[0], self.k, nodes=inp[1])

    def reference_loss(self, ref, net, inp):
        inp[0].requires_grad_(True)          # fredholm.py:101
        return ref.fredholm.dgm_loss_func(net, inp[0], self.k)   # draws its own k node sets (rand_like)

    def port_loss(self, rp):
        return rp.fredholm_loss


WORKLOADS = {"heat": Heat, "ode": Ode, "fhn": Fhn, "fredholm": Fredholm}


def slice_rows(wl, inp, n):
    if wl.name == "fredholm":
        return [inp[0][:n].contiguous(), inp[1][:, :n].contiguous()]
    return [z[:n].contiguous() for z in inp]


# ------------------------------------------------------------------ the reference, executed
def load_reference():
    """(namespace, kind): the unmodified reference from oracle/_ref ('reference') or None ('port')."""
    from oracle import ref_loader
    if ref_loader.available():
        with quiet():
            return ref_loader.load(), "reference"
    return None, "port"


def reference_runner(wl, device, B, seed=1):
    """-> (step() -> float loss, kind).  One step = zero_grad + the reference's dgm_loss_func + backward +
    torch.optim.Adam.step + loss.item(): the body of the reference's training loop (heat.py:136-143)."""
    ref, kind = load_reference()
    inp = [z.to(device) for z in slice_rows(wl, wl.make_inputs(B, seed), B)]
    if ref is not None:
        net = wl.build_net(ref).to(device)
        opt = torch.optim.Adam(net.parameters(), lr=1e-4)

        def step():
            opt.zero_grad()
            loss = wl.reference_loss(ref, net, inp)
            loss.backward()
            opt.step()
            return loss.item()
        return step, kind
    # oracle/_ref absent (the recipe never ran): the restatement of the same algorithm
    from oracle import ref_port as rp
    kinds = {"mlp": rp.KIND_MLP, "dgm": rp.KIND_DGM_LINEAR, "dgmraw": rp.KIND_DGM_RAW}
    spec = rp.NetSpec(kinds[wl.kind], wl.d, wl.o, wl.H, wl.L, rp.ACT_NAMES[wl.act])
    gen = torch.Generator().manual_seed(1234)
    state = {"theta": ((torch.rand(spec.num_params(), generator=gen) - 0.5) * 0.2).to(device)}
    state["m"], state["v"], state["t"] = torch.zeros_like(state["theta"]), torch.zeros_like(state["theta"]), 0
    fn = wl.port_loss(rp)

    def step():
        loss, g = rp.loss_and_grad(fn, spec, state["theta"], *inp)
        state["t"] += 1
        state["theta"], state["m"], state["v"] = rp.adam_step(state["theta"], state["m"], state["v"], g, state["t"])
        return float(loss)
    return step, kind


def cpu_reference(wl, B, steps, warmup, budget_s):
    """The reference on the host cores.  The sample is min(B, 2^16) rows, reduced by
    powers of two only as far as needed for `warmup + steps` steps to fit `budget_s` (rows/s is flat beyond
    ~4096 rows: BASELINE.md section 2)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    floor = min(B, 256 if wl.name == "fredholm" else 4096)
    probe_step, kind = reference_runner(wl, "cpu", floor)
    probe_step()
    t0 = time.perf_counter()
    probe_step()
    rate = floor / (time.perf_counter() - t0)
    n = min(B, 1 << 16)
    while n > floor and n * (steps + warmup) / rate > budget_s:
        n //= 2
    step, kind = reference_runner(wl, "cpu", n) if n != floor else (probe_step, kind)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    what = "unmodified reference (oracle/_ref: dgm_loss_func + backward + torch.optim.Adam.step)" if kind == "reference" \
        else "oracle/ref_port (restatement of the reference algorithm) + Adam"
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{n} rows/step of the same workload, {len(times)} timed steps after {warmup} warm-up, {what}, "
                      f"torch {torch.__version__} CPU, {cores} threads", "ms_per_step": dt * 1e3, "rows": n}


def cuda_eager_reference(wl, B, dev):
    """SURVEY 8(d) / BASELINE.md section 3, second comparator: the same reference code as eager torch on THIS
    B200 (cuBLAS SGEMM FP32, TF32 off), CUDA events, 3 warm-ups -- the pre-existing Blackwell path to beat."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    n = min(B, 1 << 16) if wl.H >= 64 else min(B, 1 << 18)
    if wl.name == "fredholm":
        n = min(B, 1 << 11)
    step, kind = reference_runner(wl, dev, n)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"value": n / (ms * 1e-3), "unit": UNIT, "kind": kind, "ms_per_step": ms, "rows": n,
            "what": f"eager torch {torch.__version__} on the same GPU, FP32 cuBLAS (TF32 off), {n} rows/step, "
                    f"step = dgm_loss_func + backward + torch.optim.Adam.step + loss.item(), {reps} timed steps after 3 warm-ups"}


def make_config(wl, B, world, scaling):
    return {"workload": f"{wl.file} loss + {wl.net_name()}, {B} rows/GPU ({wl.tag})",
            "net": wl.net_name(), "rows_per_gpu": B, "global_rows": B * max(world, 1), "parallelism": f"dp{max(world, 1)}",
            "step": "fused loss+grad kernels, all-reduce(grad|loss) if N>1, fused Adam",
            "arithmetic": "FP32 in/out; hidden size 128: GEMMs on tcgen05 with 3xTF32 split + RN chunk accumulation "
                          "(measured 1.2e-7 vs FP64, FP32 FFMA tile: 2.0e-7); otherwise FP32 FFMA; element-wise jets in FP32",
            "l2": "per-step working set (activation stash, GBs per chunk) >> 126 MB L2 at hidden size 128; "
                  "inputs larger than L2 are re-read from HBM every step"}


def main():
    a = ARGS
    wl = WORKLOADS[a.config](a)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if a.scaling == "strong":
        G = a.global_rows or (1 << 24)
        B = G // max(world, 1)
    else:
        B = a.rows_per_gpu or wl.default_rows
    config = make_config(wl, B, world, a.scaling)

    if a.impl == "reference":
        if rank != 0:
            return
        W = max(a.warmup, 1)
        r = cpu_reference(wl, B, a.steps, W, budget_s=150.0)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": W, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": a.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}), flush=True)
        return

    import ctypes as C
    import types
    import torch.distributed as dist
    from differential_equations_dnn_b200 import (_cabi, dgm_net, neural_networks, heat, simple_ode, fitzhugh_nagumo,
                                                 fredholm, optim, parallel, _loop)
    pk = types.SimpleNamespace(dgm_net=dgm_net, neural_networks=neural_networks, heat=heat, simple_ode=simple_ode,
                               fitzhugh_nagumo=fitzhugh_nagumo, fredholm=fredholm, _loop=_loop)

    lib = _cabi.load()  # raises if the CUDA library is missing: no fallback
    assert torch.cuda.is_available(), "bench.py (b200 arm) needs a GPU"
    # stdout carries exactly one JSON line: keep NCCL's own banner out of it
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    parallel.init_from_env("nccl")
    if dist.is_initialized():
        parallel.enable_data_parallel()
        world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", torch.cuda.current_device())
    W = max(a.warmup, 3)

    net = wl.build_net(pk).to(dev)
    opt = optim.FusedAdam(net.parameters(), lr=1e-4)
    host = [t.pin_memory() for t in wl.make_inputs(B, 1 + rank)]
    res = [t.to(dev) for t in host]
    stage = [torch.empty_like(t) for t in res]

    def barrier():
        if dist.is_initialized():
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        opt.zero_grad()
        loss = wl.product_loss(pk, net, res)
        loss.backward()
        opt.step()
        return loss

    def step_e2e():
        for d_, h_ in zip(stage, host):
            d_.copy_(h_, non_blocking=True)   # H2D of this step's rows from pinned memory
        opt.zero_grad()
        loss = wl.product_loss(pk, net, stage)
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

    ms, wall_ms, clocks, launches, last = timed(step_resident, W, a.steps, True)
    ms_step = ms / a.steps
    rows = B * max(world, 1)
    value = rows / (ms_step * 1e-3)
    # per-class timing pass: CUDA events around every kernel class of the library, on the launch stream -- a
    # SEPARATE pass after the timed region, so the events cost nothing inside it
    lib.dgmk_profile(1)
    PROF_STEPS = max(2, min(a.steps, 5))
    for _ in range(PROF_STEPS):
        step_resident()
    torch.cuda.synchronize()
    lib.dgmk_profile(0)
    prof = {}
    names = {0: "wg::wgrad_ws_kernel (weight gradient: warp-specialised tcgen05 kind::tf32, A^T in tensor memory, 3xTF32)",
             1: "lg::lane_gemm_kernel (fused GEMM + jet stage: weights in tensor memory, warp-specialised, 3xTF32)",
             2: "dg::dgrad_res_kernel (K = 3H data gradient: weights resident in shared memory, rows by TMA through tensor memory, 3xTF32; FFMA2 tiles for H % 128 != 0)",
             3: "ew_kernel<...> / rev1_ev_kernel (stand-alone element-wise jet stages, loss, pack, Adam)",
             4: "wcolsum / rowdot / reduce_partials (output layer, column sums, second-stage reductions)",
             5: "tk::tile_step_kernel (hidden size <= 64: resident-tile step -- one persistent kernel, a tile of points through "
                "forward jets, loss and reverse with stash and packed weights in shared memory, FP32 FFMA2, per-CTA "
                "gradient accumulators)"}
    for cls in range(6):
        t_, n_, f_, b_ = C.c_double(), C.c_longlong(), C.c_double(), C.c_double()
        if lib.dgmk_profile_read(cls, C.byref(t_), C.byref(n_), C.byref(f_), C.byref(b_)) != 0:
            continue
        if n_.value:
            prof[cls] = {"kernel": names[cls], "ms_per_step": t_.value / PROF_STEPS, "launches_per_step": n_.value / PROF_STEPS,
                         "alg_flops_per_step": f_.value / PROF_STEPS, "design_bytes_per_step": b_.value / PROF_STEPS}
    e_ms, e_wall, _, _, _ = timed(step_e2e, 2, a.steps, False)
    e_step = max(e_ms, e_wall) / a.steps      # host-side copies/sync: take the larger clock
    e2e_value = rows / (e_step * 1e-3)

    if rank != 0:
        if dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()
        return
    roof = roofline(lib, wl, prof, B, ms_step, dev)

    cpu = cuda_eager = None
    if world == 1 and not a.no_cpu_baseline:
        try:
            r = cpu_reference(wl, B, 3, 1, budget_s=25.0)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:
            cpu = {"error": repr(e)}
    if world == 1 and not a.no_cuda_eager:
        try:
            cuda_eager = cuda_eager_reference(wl, B, dev)
        except Exception as e:
            cuda_eager = {"error": repr(e)}

    drv = driver_latency(pk, wl, dev) if (world == 1 and not a.no_driver_latency) else None
    h2d = sum(t.numel() * 4 for t in host)
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": max(world, 1), "steps": a.steps, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": config, "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": e_step},
        "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "cuda_eager_baseline": cuda_eager,
        "driver_latency": drv,
        "node_evals_per_sec": value * (wl.k + 1) if wl.name == "fredholm" else None,
        "loss_last": float(last.detach()) if hasattr(last, "detach") else float(last), "wall_ms_per_step": wall_ms / a.steps}), flush=True)
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


def driver_latency(pk, wl, dev):
    """The reference's own regime: its training driver at the SHIPPED batch size (heat.py / simple_ode.py:
    64 rows, fitzhugh_nagumo.py: 100, fredholm.py: 32 rows x k = 50) is launch-bound.  us per iteration of this package's
    `minimize_loss_dgm` for the workload's network -- sampler, fused step, Adam, loss record -- launched from Python
    (eager; difference of two run lengths) and replayed from a CUDA graph (cuda_graph=True; CUDA events around the replays)."""
    import contextlib
    import io
    rows = {"heat": 64, "ode": 64, "fhn": 100, "fredholm": 32}[wl.name]

    def run(n, graph, philox=False):
        smp = {"sampler": "philox"} if philox else {}
        net = wl.build_net(pk).to(dev)
        with contextlib.redirect_stdout(io.StringIO()):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if wl.name == "heat":
                pk.heat.minimize_loss_dgm(net, iterations=n, batch_size=rows, lrate=1e-4, cuda_graph=graph, **smp)
            elif wl.name == "ode":
                pk.simple_ode.minimize_loss_dgm(net, y_ic=2.0, iterations=n, batch_size=rows, lrate=1e-4, cuda_graph=graph, **smp)
            elif wl.name == "fhn":
                pk.fitzhugh_nagumo.minimize_loss_dgm(net, torch.zeros([rows, 2], device=dev), iterations=n, batch_size=rows,
                                                     lrate=1e-4, sampler="philox" if philox else "grid", cuda_graph=graph)
            else:
                pk.fredholm.minimize_loss_dgm(net, iterations=n, batch_size=rows, lrate=1e-4, k=50, cuda_graph=graph, **smp)
            torch.cuda.synchronize()
        return time.perf_counter() - t0
    try:
        run(30, False), run(30, True)   # warm the caches both modes rely on
        eager = (run(330, False) - run(30, False)) / 300
        run(1011, True)                 # CUDA events around the 1000 replays (_loop.last_timing): capture excluded
        lt = pk._loop.last_timing
        graph = lt["ms"] * 1e-3 / max(lt["replays"], 1)
        run(30, True, True)
        run(1011, True, True)           # the same with the on-device Philox sampler (one sampler launch per iteration)
        lt = pk._loop.last_timing
        graph_philox = lt["ms"] * 1e-3 / max(lt["replays"], 1)
        return {"rows": rows, "k": 50 if wl.name == "fredholm" else None, "eager_us_per_iteration": eager * 1e6,
                "cuda_graph_us_per_iteration": graph * 1e6, "cuda_graph_philox_us_per_iteration": graph_philox * 1e6,
                "what": "this package's minimize_loss_dgm at the reference driver's shipped batch size (sampler + fused step + "
                        "fused Adam + loss record per iteration); eager: difference of two run lengths, graph: CUDA events around 1000 replays"}
    except Exception as e:
        return {"error": repr(e)}


def roofline(lib, wl, prof, B, ms_step, dev):
    """Roofline of the dominant kernel class, from CUDA events around its launches (dgmk_profile pass).

    What is what:
      alg_flops          the class's share of F_alg = 2*M*N*K of its contractions (element-wise work not counted)
      design_bytes       what THIS design moves through HBM for the class: every operand read / result written once,
                         activation stash included.  NOT algorithmic: the path's algorithmic HBM input is
                         `alg_input_bytes_per_row` (40 B per heat row); the stash is a design choice.
      bound              "tensor": the contractions run on the tensor pipe (kind::tf32, three MMAs per FP32-grade
                         product); peak = dense TF32 = 1/2 x the measured bf16 figure.  frac = alg TFLOP/s / peak; the
                         kernel's own ceiling is a third of that (`frac_of_3xtf32_ceiling`).  `design_hbm_frac` says how
                         close the class runs to the HBM roofline of the bytes the design moves.
                         "fp32": hidden sizes that are not 128 run on the FP32 FMA pipe; peak = live FFMA probe.
    """
    import ctypes as C
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
        hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6500 (B200_PROFILING.md)"
        tf32_peak = bf16 / 2.0
        classes = {}
        for cls, p in prof.items():
            t = p["ms_per_step"] * 1e-3
            classes[str(cls)] = dict(p, share_of_step=p["ms_per_step"] / ms_step,
                                     alg_tflops=p["alg_flops_per_step"] / t / 1e12 if p["alg_flops_per_step"] else None,
                                     design_gbs=p["design_bytes_per_step"] / t / 1e9 if p["design_bytes_per_step"] else None)
        gemm_cls = [c for c in prof if prof[c]["alg_flops_per_step"]]
        dom = max(gemm_cls, key=lambda c: prof[c]["ms_per_step"])
        k = classes[str(dom)]
        per_launch = 1.0 / k["launches_per_step"]
        on_tensor = dom in (0, 1) or (dom == 2 and wl.H % 128 == 0)
        step_bytes = sum(p["design_bytes_per_step"] for p in prof.values())
        alg_in = wl.in_bytes * B
        fl_row = wl.f_alg()
        roof = {"kernel": k["kernel"],
                "launch": {"avg_ms": k["ms_per_step"] * per_launch, "alg_flops": k["alg_flops_per_step"] * per_launch,
                           "design_bytes": k["design_bytes_per_step"] * per_launch, "per_step": k["launches_per_step"]},
                "share_of_step": k["share_of_step"],
                "design_hbm_gbs": k["design_gbs"], "hbm_peak_gbs": hbm, "hbm_peak_source": hbm_src,
                "design_hbm_frac": (k["design_gbs"] or 0.0) / hbm,
                # whole step: what the design moves vs what the path needs (40 B per heat row)
                "alg_input_bytes_per_row": wl.in_bytes, "alg_input_bytes_per_step": alg_in,
                "step_design_bytes_moved": step_bytes, "step_design_bytes_over_alg_input": step_bytes / alg_in,
                "step_hbm_floor_ms": step_bytes / (hbm * 1e9) * 1e3, "step_hbm_floor_share": step_bytes / (hbm * 1e9) * 1e3 / ms_step,
                "kernel_classes": classes, "traffic": None,
                "fp32_ffma_peak_live": fp32_peak, "alg_flops_per_row": fl_row,
                "step_alg_tflops": fl_row * B / (ms_step * 1e-3) / 1e12,
                "step_frac_of_fp32_peak": fl_row * B / (ms_step * 1e-3) / 1e12 / fp32_peak,
                "step_frac_of_tf32_peak": fl_row * B / (ms_step * 1e-3) / 1e12 / tf32_peak,
                "step_frac_of_3xtf32_ceiling": 3 * fl_row * B / (ms_step * 1e-3) / 1e12 / tf32_peak}
        if on_tensor:
            roof.update(bound="tensor", achieved=k["alg_tflops"], peak=tf32_peak, unit="TFLOP/s", frac=k["alg_tflops"] / tf32_peak,
                        frac_of_3xtf32_ceiling=3 * k["alg_tflops"] / tf32_peak,
                        peak_source=f"TF32 dense = 1/2 x {bf16_src} = {tf32_peak:.1f} TFLOP/s; achieved = algorithmic 2MNK flops of "
                                    "the class's launches / their CUDA-event durations (the kernel issues 3 TF32 MMAs per "
                                    "FP32-grade product: its own ceiling is peak/3)")
        else:
            roof.update(bound="fp32", achieved=k["alg_tflops"], peak=fp32_peak, unit="TFLOP/s", frac=k["alg_tflops"] / fp32_peak,
                        peak_source="FP32 FFMA peak measured live (dgmk_ffma_probe: dependent FFMA chains, 16 per thread); "
                                    "achieved = algorithmic 2MNK flops of the class's launches / their CUDA-event durations")
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            tj = json.load(open(tr))
            key = {0: "wgrad_ws", 1: "lane_gemm", 2: "dgrad_res", 5: "tile_step"}.get(dom)
            ratio = tj.get(key + "_dram_over_algorithmic") if key else None
            if ratio is not None:   # ncu dram__bytes_read+write per launch / design bytes of that launch
                roof["traffic"] = ratio * roof["launch"]["design_bytes"]
                roof["traffic_source"] = tj.get(key + "_source")
        return roof
    except Exception as e:  # the headline number is still valid without the probe
        return {"error": repr(e)}


if __name__ == "__main__":
    main()
