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
y, s, n = make_batch(B, world, C, L, seed0=21)
yd, sd, nd = (torch.from_numpy(a).to(dev) for a in (y, s, n))
# all nodes on this GPU, oracle masks (informative masks keep the GEVD well conditioned: with random masks R_ss ~ c R_nn
# and the principal generalised eigenvector amplifies the last-bit differences between two summation orders)
full = tango_batched(yd, sd, nd, out_layout="TF", diagnostics=False)
mzd, mwd = full["masks_z"].contiguous(), full["mask_w"].contiguous()
res = tango_node_sharded(yd[:, rank:rank + 1].contiguous(), mzd[:, rank:rank + 1].contiguous(),
                         mwd[:, rank:rank + 1].contiguous())
torch.cuda.synchronize()
rel = lambda a, b: (torch.linalg.norm(a - b) / torch.linalg.norm(b)).item()
e_yf = rel(res["yf"][:, 0].abs(), full["yf"][:, rank].abs())
e_z = rel(res["Z"].abs(), full["z_y"].abs())
print("rank %d/%d node-sharded (NCCL all-gather of z) vs single-GPU: |yf| rel-L2 %.2e, |Z| rel-L2 %.2e"
      % (rank, world, e_yf, e_z), flush=True)
assert e_yf < 1e-5 and e_z < 1e-5          # the parity tolerance of the path (the two routes sum partial SCMs in different orders)
dist.destroy_process_group()
