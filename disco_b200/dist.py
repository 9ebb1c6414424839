"""Multi-GPU execution of the Tango path: one process per GPU, torch.distributed for the plumbing.

Two ways to shard:
* utterance sharding -- utterances are independent, rank r takes its slice and runs
  ``tango_batched`` locally; NO data-path collective (``shard_range``);
* node sharding -- rank r owns array node r of every utterance (the physical layout of a
  distributed microphone array).  The reference's in-process exchange of the compressed signals
  (tango.py:379-386: every node receives the z of all the others) becomes ONE all-gather of
  z [B, T, F] complex64 between step 1 and step 2 (``tango_node_sharded``).

The compute callables are injectable so that the exchange logic (gather order, own-node indexing)
can be exercised with the gloo backend on CPU-only machines (tests/test_dist_cpu.py).
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced [lo, hi) slice of n_items for `rank` (sizes differ by at most one)."""
    return (n_items * rank) // world, (n_items * (rank + 1)) // world


def all_gather_nodes(z_local, group=None):
    """z_local [B, T, F] of this rank's node -> Z [B, K, T, F] with node k = rank k's tensor."""
    world = dist.get_world_size(group)
    z_local = z_local.contiguous()
    flat = torch.view_as_real(z_local) if z_local.is_complex() else z_local
    # concatenated form [K * B, ...] (accepted by both NCCL and gloo), viewed as [K, B, ...]
    out = torch.empty((world * flat.shape[0],) + tuple(flat.shape[1:]), dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(out, flat, group=group)
    out = out.view((world,) + tuple(flat.shape))
    Z = torch.view_as_complex(out) if z_local.is_complex() else out
    return Z.transpose(0, 1).contiguous()          # [K, B, ...] -> [B, K, ...]


def _gpu_step1(y, mask_z, **kw):
    from .tango import tango_step1
    return tango_step1(y, mask_z, **kw)


def _gpu_step2(Y, Z, mask_w, node, **kw):
    from .tango import tango_step2
    return tango_step2(Y, Z, mask_w, node_sel=[node], **kw)[0]


def tango_node_sharded(y_local, mask_z, mask_w=None, group=None, step1=_gpu_step1, step2=_gpu_step2, **kw):
    """Two-step Tango with one array node per rank.

    y_local [B, 1, C, L] -- this rank's node; mask_z / mask_w [B, 1, T, F].
    Returns dict(yf [B, 1, T, F], z_y [B, 1, T, F], Z [B, K, T, F])."""
    rank = dist.get_rank(group)
    st1 = step1(y_local, mask_z, **kw)
    z_local = st1["z_y"][:, 0]                      # [B, T, F]
    Z = all_gather_nodes(z_local, group)            # the exchange step (NCCL over NVLink on GPUs)
    yf = step2(st1["Y"], Z, mask_z if mask_w is None else mask_w, rank, **kw)
    return {"yf": yf, "z_y": st1["z_y"], "Z": Z}
