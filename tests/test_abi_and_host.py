"""CPU-only checks: the C-ABI library loads and exports every symbol include/dgmk.h
declares (no compute without a GPU), layouts agree between Python, the C ABI and the
oracle, and the drop-in modules keep the reference's contract."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, golden, golden_names


@pytest.fixture(scope="module")
def cabi():
    from differential_equations_dnn_b200 import _cabi
    _cabi.build()   # nvcc cross-compiles sm_100a without a GPU
    return _cabi


def test_library_exports_every_declared_symbol(cabi):
    hdr = open(os.path.join(ROOT, "include", "dgmk.h")).read()
    declared = set(re.findall(r"\b(dgmk_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"dgmk_net_desc"}
    assert declared, "no declarations parsed"
    lib = C.CDLL(cabi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/dgmk.h but not exported"
    assert declared == set(cabi.EXPORTS), declared ^ set(cabi.EXPORTS)
    lib = cabi.load()
    assert lib.dgmk_version() == 100 and lib.dgmk_backend() == b"cuda-sm100a"


def test_sass_is_sm100a(cabi):
    """The built library is sm_100a SASS and its contraction kernels are Blackwell-native: tcgen05 MMAs
    (UTCHMMA), tensor-memory loads / stores (LDTM / STTM), the bulk-copy engine (UBLKCP), tcgen05.commit (UTCBAR)
    -- the histogram tools/sass_hist.py writes to profiles/r02_sass_histogram.txt."""
    import subprocess
    import sys
    out = subprocess.run(["cuobjdump", "-lelf", cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_hist
    h = sass_hist.histogram(cabi.LIB_PATH)
    tot = {}
    for c in h.values():
        for op, n in c.items():
            tot[op] = tot.get(op, 0) + n
    for op in ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "FFMA2"):
        assert tot.get(op, 0) > 0, f"no {op} in the library's SASS"
    for name, c in h.items():
        if "lane_gemm_kernel" in name or "wgrad_ws_kernel" in name:
            assert c["UTCHMMA"] >= 12 and c["LDTM"] > 0 and c["UBLKCP"] + c["LDG"] > 0, name
        if "lane_gemm_kernel" in name:
            assert c["STTM"] > 0 and c["UBLKCP"] > 0, name     # weights resident in tensor memory, rows by bulk copy
        if "dgrad_res_kernel" in name:   # rows by TMA tensor copies into tensor memory, weights resident in shared memory
            assert c["UTCHMMA"] >= 12 and c["UTMALDG"] > 0 and c["STTM"] > 0 and c["LDTM"] > 0, name
        if "small_step_kernel" in name:
            assert c["UBLKCP"] > 0 and c["FFMA"] > 100, name   # weights staged by the bulk-copy engine, FP32 FMA pipe


def test_missing_library_fails_loudly(cabi, monkeypatch):
    monkeypatch.setattr(cabi, "_LIB", None)
    monkeypatch.setattr(cabi, "LIB_PATH", "/nonexistent/libdgmk.so")
    with pytest.raises(cabi.DgmkError, match="no CPU or PyTorch fallback"):
        cabi.load()


@pytest.mark.parametrize("name", [n for p in ("heat_", "ode_", "fhn_", "fredholm_") for n in golden_names(p)
                                  if "driver" not in n])
def test_layouts_agree(cabi, name):
    from oracle import ref_port as rp
    from differential_equations_dnn_b200 import kernels
    g = golden(name)
    spec = rp.NetSpec(*[int(v) for v in g["spec"]])
    d = cabi.make_desc(*[int(v) for v in g["spec"]])
    assert kernels.param_count(d) == spec.num_params() == g["theta"].size
    lay = kernels.param_layout(d)
    off = 0
    for (o, r, c, live), (nm, shape, olive) in zip(lay, spec.entries()):
        assert o == off and ((r,) if c == 0 else (r, c)) == tuple(shape) and live == olive, nm
        off += int(np.prod(shape))
    assert len(lay) == len(spec.entries())
    assert kernels.workspace_bytes(d, cabi.WS_HEAT if spec.d == 2 else cabi.WS_ODE if spec.o == 1 else cabi.WS_FHN, 1024) > 0


def test_same_seed_same_weights_as_reference():
    from differential_equations_dnn_b200 import neural_networks as nn_, dgm_net
    cases = {"heat_dgm_h128l3": lambda: dgm_net.DGM(2, 1, 128, 3),
             "heat_mlp_relu_h128l3": lambda: nn_.MLP(2, 1, 128, 3),
             "heat_mlp_sigmoid_h50l1": lambda: nn_.MLP(2, 1, 50, 1, activation="sigmoid"),
             "heat_mlp_leaky_h32l2": lambda: nn_.MLP(2, 1, 32, 2, activation="leaky_relu"),
             "fhn_mlp_tanh_h128l3": lambda: nn_.MLP(1, 2, 128, 3, activation="tanh"),
             "fredholm_dgmraw_h32l1_k50": lambda: nn_.DGM(1, 1, 32, 1)}
    for name, ctor in cases.items():
        torch.manual_seed(1234)
        net = ctor()
        assert np.array_equal(net.flat_theta().numpy(), golden(name)["theta"]), name


def test_module_contract_cpu():
    from differential_equations_dnn_b200 import neural_networks as nn_, dgm_net
    net = nn_.DGM(input_dim=1, output_dim=1, hidden_size=32)
    names = [n for n, _ in net.named_parameters()]
    assert names[:3] == ["x_in.weight", "x_in.bias", "dgm1.Uz"] and names[-1] == "x_out.bias"
    assert tuple(net.layers[0].Uz.shape) == (1, 32) and tuple(net.layers[0].bz.shape) == (1, 32)
    assert int(net.live_mask().sum()) == 4449 and net.flat_theta().numel() == 8801
    # parameters are views of one buffer, and stay so after dtype/device style conversions
    f = net.flat_theta()
    with torch.no_grad():
        net.x_out.bias.fill_(3.0)
    assert f[-1].item() == 3.0
    net = net.float()
    f2 = net.flat_theta()
    with torch.no_grad():
        net.x_in.weight.zero_()
    assert f2[:32].abs().sum().item() == 0
    m = dgm_net.DGM(2, 1, 128, 3)
    assert m.flat_theta().numel() == 201729
    assert [k for k in m.state_dict()][2] == "layers.0.Z_wg.weight"
    with pytest.raises(NotImplementedError):
        nn_.MLP(batch_norm=True)
    with pytest.raises(Exception, match="no CPU"):
        m(torch.zeros(4, 2))


def test_fused_adam_binds_to_flat_buffer():
    from differential_equations_dnn_b200 import dgm_net
    from differential_equations_dnn_b200.optim import FusedAdam
    net = dgm_net.DGM(1, 1, 8, 1)
    opt = FusedAdam(net.parameters(), lr=1e-4)
    assert opt.param_groups[0]["lr"] == 1e-4 and opt.net is net
    with pytest.raises(ValueError):
        FusedAdam([torch.nn.Parameter(torch.zeros(3))])


def test_search_space_and_trials():
    from differential_equations_dnn_b200 import parallel
    cfgs = parallel.sample_search_space(10, seed=0)
    assert len(cfgs) == 10 and all(1 <= c["batch_size"] < 512 and 1000 <= c["n_iters"] < 50000
                                   and 1e-4 <= c["lrate"] <= 1e-1 for c in cfgs)
    res = parallel.run_trials(lambda c: c["lrate"], cfgs)
    assert parallel.best_trial(res)["loss"] == min(c["lrate"] for c in cfgs)


def test_modules_pickle_deepcopy_and_optimizer_state():
    """The drop-in modules support what the reference's do: torch.save(net) / pickle / copy.deepcopy (nothing
    unpicklable hangs on the Parameters), the copy is re-tied to its own flat buffer, FusedAdam accepts its
    parameters, and the optimizer's state_dict carries the flat moments and the step count."""
    import copy
    import io
    import pickle
    import torch
    from differential_equations_dnn_b200 import dgm_net, neural_networks, optim
    for net in (dgm_net.DGM(2, 1, 16, 2), neural_networks.MLP(1, 2, 8, 1, activation="tanh"),
                neural_networks.DGM(1, 1, 8, 1, func="tanh")):
        ref = net.flat_theta().clone()
        buf = io.BytesIO()
        torch.save(net, buf)
        buf.seek(0)
        for c in (copy.deepcopy(net), pickle.loads(pickle.dumps(net)), torch.load(buf, weights_only=False)):
            assert torch.equal(c.flat_theta(), ref) and c.flat_theta().data_ptr() != net.flat_theta().data_ptr()
            with torch.no_grad():
                c.flat_theta().add_(1.0)        # parameters are views of the copy's own buffer
            assert all(torch.equal(p.detach().reshape(-1), c.flat_theta()[off:off + n]) for p, off, n, _ in c.param_slices())
            assert torch.equal(net.flat_theta(), ref)
            opt = optim.FusedAdam(c.parameters(), lr=1e-3)
            assert opt.net is c
        opt = optim.FusedAdam(net.parameters(), lr=1e-3)
        opt._t, opt._m, opt._v = 7, torch.full_like(ref, 0.5), torch.full_like(ref, 0.25)
        sd = opt.state_dict()
        opt2 = optim.FusedAdam(net.parameters(), lr=1e-2)
        opt2.load_state_dict(sd)
        assert opt2._t == 7 and torch.equal(opt2._m, opt._m) and torch.equal(opt2._v, opt._v)
        assert opt2.param_groups[0]["lr"] == 1e-3
    with pytest.raises(ValueError):
        optim.FusedAdam(torch.nn.Linear(2, 2).parameters())
