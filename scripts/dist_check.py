"""torchrun --nproc-per-node K scripts/dist_check.py : node-sharded Tango over NCCL == single-GPU result."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from disco_b200.synth import make_batch
from disco_b200.tango import tango_batched
from disco_b200.dist import tango_node_sharded

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
B, C, L = 4, 4, 32000
y, _, _ = make_batch(B, world, C, L, seed0=21)
T, F = 1 + L // 256, 257
g = torch.Generator().manual_seed(5)
mz = torch.rand((B, world, T, F), generator=g)
mw = torch.rand((B, world, T, F), generator=g)
yd, mzd, mwd = torch.from_numpy(y).to(dev), mz.to(dev), mw.to(dev)
full = tango_batched(yd, masks=(mzd, mwd), out_layout="TF", diagnostics=False)          # all nodes on this GPU
res = tango_node_sharded(yd[:, rank:rank + 1].contiguous(), mzd[:, rank:rank + 1].contiguous(),
                         mwd[:, rank:rank + 1].contiguous())
torch.cuda.synchronize()
e_yf = (res["yf"][:, 0] - full["yf"][:, rank]).abs().max().item() / full["yf"].abs().max().item()
e_z = (res["Z"] - full["z_y"]).abs().max().item()
print("rank %d/%d node-sharded vs single-GPU: yf rel-max %.2e, Z abs-max %.2e" % (rank, world, e_yf, e_z), flush=True)
assert e_yf < 1e-6 and e_z == 0.0
dist.destroy_process_group()
