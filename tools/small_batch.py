"""Per-step latency of the training driver at the reference's batch sizes (launch-bound regime):
eager loop vs the CUDA-graph replay (heat.minimize_loss_dgm(cuda_graph=True))."""
import sys, time, torch
sys.path.insert(0, ".")
from differential_equations_dnn_b200 import dgm_net, heat
torch.manual_seed(0)
for H, L in ((128, 3), (32, 1)):
    for B in (64, 256, 4096):
        row = []
        for graph in (False, True):
            net = dgm_net.DGM(2, 1, H, L).cuda()
            heat.minimize_loss_dgm(net, iterations=100, batch_size=B, lrate=1e-4, cuda_graph=graph)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            heat.minimize_loss_dgm(net, iterations=2000, batch_size=B, lrate=1e-4, cuda_graph=graph)
            torch.cuda.synchronize()
            row.append((time.perf_counter() - t1) / 2000 * 1e3)
        print(f"DGM(2,1,{H},{L}) B={B}: eager {row[0]:.3f} ms/step   cuda_graph {row[1]:.3f} ms/step   ({row[0]/row[1]:.1f}x)", flush=True)

# the other three drivers at the reference's shipped sizes (simple_ode.py:157-167, fitzhugh_nagumo.py:202-214,
# fredholm.py:160-173)
import io, contextlib
from differential_equations_dnn_b200 import neural_networks, simple_ode, fitzhugh_nagumo, fredholm
def timed(fn, its=1000):
    with contextlib.redirect_stdout(io.StringIO()):
        fn(100)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        fn(its)
        torch.cuda.synchronize()
    return (time.perf_counter() - t1) / its * 1e3
for name, mk, run in (
    ("simple_ode MLP(1,1,32), B=64", lambda: neural_networks.MLP(input_dim=1, output_dim=1, hidden_size=32).cuda(),
     lambda net, g: (lambda n: simple_ode.minimize_loss_dgm(net, y_ic=2.0, iterations=n, batch_size=64, lrate=1e-4, cuda_graph=g))),
    ("fhn dgm_net.DGM(1,2,128,4), B=100, uniform sampler", lambda: dgm_net.DGM(1, 2, 128, 4).cuda(),
     lambda net, g: (lambda n: fitzhugh_nagumo.minimize_loss_dgm(net, torch.zeros([100, 2], device="cuda"), iterations=n, batch_size=100,
                                                               lrate=1e-4, sampler="uniform", cuda_graph=g))),
    ("fredholm neural_networks.DGM(1,1,32), B=32, k=50", lambda: neural_networks.DGM(input_dim=1, output_dim=1, hidden_size=32).cuda(),
     lambda net, g: (lambda n: fredholm.minimize_loss_dgm(net, iterations=n, batch_size=32, lrate=1e-4, k=50, cuda_graph=g)))):
    row = [timed(run(mk(), g)) for g in (False, True)]
    print(f"{name}: eager {row[0]:.3f} ms/step   cuda_graph {row[1]:.3f} ms/step   ({row[0]/row[1]:.1f}x)", flush=True)
