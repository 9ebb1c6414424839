"""Multi-GPU execution of the Tango path: one process per GPU, torch.distributed for the plumbing.

Two ways to shard:
* utterance sharding -- utterances are independent, rank r takes its slice and runs
  ``tango_batched`` locally; NO data-path collective (``shard_range``);
* node sharding -- rank r owns K / world array nodes of every utterance (the physical layout of a
  distributed microphone array).  The reference's in-process exchange of the compressed signals
  (tango.py:379-386: every node receives the z of all the others) becomes an all-gather of
  z [B, T, F] complex64 between step 1 and step 2 (``tango_node_sharded``):
    - the batch is cut into chunks; the all-gather of chunk i runs on a communication stream while
      step 1 of chunk i + 1 computes, and step 2 of chunk i starts as soon as its gather has landed;
    - the gather lands in a node-major buffer [K, B_chunk, T, F] that the step-2 kernels index in
      place (``z_layout='KB'``): the gathered signals are never transposed or copied.

The compute callables are injectable so that the exchange logic (gather order, own-node indexing,
chunk pipeline) can be exercised with the gloo backend on CPU-only machines (tests/test_dist_cpu.py).
"""
import contextlib
import inspect

import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced [lo, hi) slice of n_items for `rank` (sizes differ by at most one)."""
    return (n_items * rank) // world, (n_items * (rank + 1)) // world


def all_gather_nodes(z_local, group=None, out=None):
    """z_local [B, Kl, T, F] (this rank's Kl nodes) -> node-major Z [K, B, T, F], K = world * Kl, node k owned
    by rank k // Kl.  With Kl = 1 the local tensor is gathered as it is (no copy on either side); with several
    nodes per rank the LOCAL z (1 / C of the size of the spectra) is put in node-major order first.
    Also accepts [B, ...] tensors of a single node per rank (returns [K, B, ...])."""
    world = dist.get_world_size(group)
    if z_local.dim() >= 4:
        B, Kl = z_local.shape[:2]
        src = z_local.transpose(0, 1)                      # [Kl, B, ...]: a view when Kl == 1
        if not src.is_contiguous():
            src = src.contiguous()
        tail = tuple(z_local.shape[2:])
    else:
        B, Kl = z_local.shape[0], 1
        src = z_local.contiguous()
        tail = tuple(z_local.shape[1:])
    flat = torch.view_as_real(src) if src.is_complex() else src
    if out is None:
        out = torch.empty((world * Kl, B) + tail, dtype=z_local.dtype, device=z_local.device)
    oflat = torch.view_as_real(out) if out.is_complex() else out
    # rank r's block [Kl, B, ...] lands at rows [r * Kl, (r + 1) * Kl) of the node-major output
    # (concatenated form: dim 0 of the output = world x dim 0 of the input, accepted by NCCL and gloo)
    dist.all_gather_into_tensor(oflat.view((world * flat.shape[0],) + tuple(flat.shape[1:])), flat, group=group)
    return out


def _filter_kw(fn, kw):
    """Keyword arguments `fn` accepts (step 1 and step 2 have different optional parameters)."""
    try:
        params = inspect.signature(fn).parameters
    except (TypeError, ValueError):
        return dict(kw)
    if any(p.kind == p.VAR_KEYWORD for p in params.values()):
        return dict(kw)
    return {k: v for k, v in kw.items() if k in params}


def _gpu_step1(y, mask_z, n_fft=512, mu=1.0, filter_type="gevd", rank=1, ref_mic=0):
    from .tango import tango_step1
    return tango_step1(y, mask_z, n_fft=n_fft, mu=mu, filter_type=filter_type, rank=rank, ref_mic=ref_mic)


def _gpu_step2(Y, Z, mask_w, nodes, n_fft=512, mu=1.0, filter_type="gevd", rank=1, out_layout="TF"):
    from .tango import tango_step2
    return tango_step2(Y, Z, mask_w, n_fft=n_fft, mu=mu, filter_type=filter_type, rank=rank, out_layout=out_layout,
                       node_sel=list(nodes), z_layout="KB")[0]


def tango_node_sharded(y_local, mask_z, mask_w=None, group=None, step1=_gpu_step1, step2=_gpu_step2, chunks=1,
                       stats=None, reserve_sms=None, **kw):
    """Two-step Tango with the array nodes sharded over the ranks of `group`.

    y_local [B, Kl, C, L] -- this rank's Kl = K / world nodes; mask_z / mask_w [B, Kl, T, F].
    chunks: batch slices of the software pipeline  step 1(i + 1)  ||  all-gather(i)  ->  step 2(i).
    kw: forwarded to step 1 / step 2, each receiving the keywords it accepts (n_fft, mu, filter_type, rank for
    both; ref_mic for step 1; out_layout for step 2).
    stats: optional dict, filled with the bytes this rank received and CUDA events around every gather.
    reserve_sms: SMs kept free of the persistent fused STFT+SCM kernel while gathers are in flight (default 16 on
    GPUs when chunks > 1, so that the NCCL kernels can run beside step 1 of the next chunk; 0 otherwise).
    Returns dict(yf [B, Kl, T, F], z_y [B, Kl, T, F], Z: list of node-major chunks [K, B_chunk, T, F])."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    B, Kl = y_local.shape[:2]
    nodes = range(rank * Kl, (rank + 1) * Kl)
    if mask_w is None:
        mask_w = mask_z
    kw1, kw2 = _filter_kw(step1, kw), _filter_kw(step2, kw)
    cuda = y_local.is_cuda
    chunks = max(1, min(int(chunks), B))
    if reserve_sms is None:
        reserve_sms = 16 if (cuda and chunks > 1 and world > 1) else 0
    if cuda and step1 is _gpu_step1:
        from . import ops
        ops.set_reserved_sms(reserve_sms)
    cuts = [shard_range(B, i, chunks) for i in range(chunks)]
    comp = torch.cuda.current_stream(y_local.device) if cuda else None
    comm = torch.cuda.Stream(device=y_local.device) if cuda else None
    on_comm = (lambda: torch.cuda.stream(comm)) if cuda else contextlib.nullcontext
    st1, Zs, done, yfs = [None] * chunks, [None] * chunks, [None] * chunks, [None] * chunks
    timing = []
    for i in range(chunks + 1):
        if i < chunks:
            lo, hi = cuts[i]
            st1[i] = step1(y_local[lo:hi], mask_z[lo:hi], **kw1)
            z_loc = st1[i]["z_y"]                                   # [Bc, Kl, T, F]
            if cuda:
                ready = torch.cuda.Event()
                ready.record(comp)
            with on_comm():
                if cuda:
                    comm.wait_event(ready)
                    z_loc.record_stream(comm)
                    if stats is not None:                           # (no timing events inside a graph capture)
                        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        t0.record(comm)
                Zs[i] = all_gather_nodes(z_loc, group)              # the exchange step (NCCL over NVLink on GPUs)
                if cuda:
                    if stats is not None:
                        t1.record(comm)
                        timing.append((t0, t1, Zs[i].numel() * Zs[i].element_size() * (world - 1) // world))
                    done[i] = torch.cuda.Event()
                    done[i].record(comm)
                    Zs[i].record_stream(comp)
        if i >= 1:
            j = i - 1
            lo, hi = cuts[j]
            if cuda:
                comp.wait_event(done[j])
            yfs[j] = step2(st1[j]["Y"], Zs[j], mask_w[lo:hi], nodes, **kw2)
    if cuda and step1 is _gpu_step1 and reserve_sms:
        ops.set_reserved_sms(0)
    if stats is not None:
        stats["gathers"] = timing
        stats["bytes_received_per_step"] = sum(t[2] for t in timing) if cuda else None
    return {"yf": torch.cat(yfs, dim=0) if chunks > 1 else yfs[0],
            "z_y": torch.cat([s["z_y"] for s in st1], dim=0) if chunks > 1 else st1[0]["z_y"], "Z": Zs}
