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
