"""torchrun --nproc-per-node N tools/dp_check.py : NCCL data-parallel step == single-process
reference on the concatenated batch (golden vector), then a few identical Adam steps."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from differential_equations_dnn_b200 import parallel, heat, dgm_net, optim

parallel.init_from_env("nccl")
parallel.enable_data_parallel()
R, r = parallel.world_size(), parallel.rank()
g = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "heat_dgm_h32l1.npz")))
torch.manual_seed(1234)
net = dgm_net.DGM(2, 1, 32, 1).cuda()
keys = ("X", "X0", "XBD1", "XBD2", "x_bd1", "x_bd2")
args = [parallel.shard(torch.from_numpy(g[k])).cuda() for k in keys]
opt = optim.FusedAdam(net.parameters(), lr=1e-3)
opt.zero_grad()
loss = heat.dgm_loss_func(net, *args)
loss.backward()
grad = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).cpu().numpy()
el = abs(loss.item() - float(g["loss"])) / abs(float(g["loss"]))
eg = np.linalg.norm(grad - g["grad"]) / np.linalg.norm(g["grad"])
assert el < 1e-5 and eg < 1e-5, (el, eg)
for _ in range(3):
    opt.step(); opt.zero_grad()
    loss = heat.dgm_loss_func(net, *args); loss.backward()
th = net.flat_theta().clone()
ref = th.clone(); dist.broadcast(ref, 0)
assert torch.equal(th, ref), "ranks diverged"
print(f"rank {r}/{R}: DP step matches the reference (loss rel {el:.1e}, grad rel {eg:.1e}); weights identical across ranks", flush=True)
dist.barrier(); dist.destroy_process_group()
