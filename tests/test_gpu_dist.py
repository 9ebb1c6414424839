"""Node-sharded distributed MWF on real GPUs (skips below 2 devices): rank r owns K / N array nodes, the compressed
signals z are all-gathered over NCCL chunk by chunk (reference tango.py:379-386) and land node-major in the buffer
the step-2 kernels read in place.  Result == all nodes on one GPU within the parity tolerance."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, K, chunks, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from dist_check import run_check
        q.put((rank,) + tuple(run_check(rank, world, rank, K=K, chunks=chunks, verbose=False)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("K,chunks", [(2, 1), (2, 2), (4, 3)])
def test_node_sharded_equals_single_gpu(K, chunks):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, K, chunks, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert [g[0] for g in got] == [0, 1] and all(g[1] < 1e-5 and g[2] < 1e-5 for g in got), got


def test_z_layout_node_major_equals_utterance_major():
    """Step-2 kernels reading Z node-major [K, B, T, F] == the same call on the transposed [B, K, T, F] buffer."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import numpy as np
    from disco_b200 import ops
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(5)
    cplx = lambda *s: torch.from_numpy((rng.standard_normal(s) + 1j * rng.standard_normal(s)).astype(np.complex64)).to(dev)
    for (B, K, C, sel) in ((3, 4, 4, [1, 2]), (2, 8, 2, [5]), (2, 2, 3, None), (2, 6, 4, [0, 3, 5])):
        T, F = 41, 257
        Ks = K if sel is None else len(sel)
        Y, Z = cplx(B, Ks, C, T, F), cplx(B, K, T, F)
        Zn = Z.transpose(0, 1).contiguous()
        m = torch.from_numpy(rng.uniform(size=(B, Ks, T, F)).astype(np.float32)).to(dev)
        D = C + K - 1
        W = cplx(B, Ks, F, D)
        Ra, Rb = ops.masked_scm(Y, m, Z, node_sel=sel)
        Rc, Rd = ops.masked_scm(Y, m, Zn, node_sel=sel, z_layout="KB")
        assert torch.equal(Ra, Rc) and torch.equal(Rb, Rd), (B, K, C)
        for lay in ("TF", "FT"):
            fa = ops.filter_sum(W, Y, Z, node_sel=sel, out_layout=lay)
            fb = ops.filter_sum(W, Y, Zn, node_sel=sel, out_layout=lay, z_layout="KB")
            if sel is None and lay == "TF":      # the all-nodes kernel reads [B, K] only; node-major takes the per-group kernel
                assert torch.allclose(torch.view_as_real(fa), torch.view_as_real(fb), rtol=1e-5, atol=1e-5)
            else:
                assert torch.equal(fa, fb), (B, K, C, lay)
