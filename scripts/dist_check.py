"""Node-sharded Tango over NCCL == all nodes on one GPU (parity of the exchange path).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/dist_check.py [K] [chunks]

Rank r owns K / N array nodes; the compressed signals are exchanged with one all-gather per batch chunk
(disco_b200/dist.py).  Also used by tests/test_gpu_dist.py (spawned workers)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def run_check(rank, world, local_rank, K=None, chunks=2, B=5, C=4, L=32000, verbose=True):
    from disco_b200.dist import tango_node_sharded
    from disco_b200.synth import make_batch
    from disco_b200.tango import tango_batched
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    K = K or world
    Kl = K // world
    assert Kl * world == K, "K must be a multiple of the number of ranks"
    y, s, n = make_batch(B, K, C, L, seed0=21)
    yd, sd, nd = (torch.from_numpy(a).to(dev) for a in (y, s, n))
    # all nodes on this GPU, oracle masks (informative masks keep the GEVD well conditioned: with random masks
    # R_ss ~ c R_nn and the principal generalised eigenvector amplifies last-bit differences between summation orders)
    full = tango_batched(yd, sd, nd, out_layout="TF", diagnostics=False)
    mzd, mwd = full["masks_z"].contiguous(), full["mask_w"].contiguous()
    sl = slice(rank * Kl, (rank + 1) * Kl)
    stats = {}
    res = tango_node_sharded(yd[:, sl].contiguous(), mzd[:, sl].contiguous(), mwd[:, sl].contiguous(), chunks=chunks,
                             stats=stats)
    torch.cuda.synchronize()
    rel = lambda a, b: (torch.linalg.norm(a - b) / torch.linalg.norm(b)).item()
    e_yf = rel(res["yf"].abs(), full["yf"][:, sl].abs())
    Z = torch.cat(res["Z"], dim=1).transpose(0, 1)                  # node-major chunks -> [B, K, T, F]
    e_z = rel(Z.abs(), full["z_y"].abs())
    if verbose:
        print("rank %d/%d nodes %d..%d, %d chunk(s): node-sharded (NCCL all-gather of z) vs single-GPU: |yf| rel-L2 %.2e, "
              "|Z| rel-L2 %.2e" % (rank, world, sl.start, sl.stop - 1, chunks, e_yf, e_z), flush=True)
    # the parity tolerance of the path (the two routes sum partial SCMs in different orders)
    assert e_yf < 1e-5 and e_z < 1e-5, (e_yf, e_z)
    return e_yf, e_z


if __name__ == "__main__":
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    K = int(sys.argv[1]) if len(sys.argv) > 1 else world
    chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    run_check(rank, world, lr, K=K, chunks=chunks)
    run_check(rank, world, lr, K=K, chunks=1, B=3, C=2, L=16000)
    dist.destroy_process_group()
