"""World-size-2 gloo tests (CPU) of the multi-GPU host logic in disco_b200/dist.py: sharding
arithmetic, all-gather order of the compressed signals, own-node indexing in step 2.  The compute
callables are replaced by the float64 oracle (the CUDA kernels need a GPU; their parity is tested
in the -m gpu suite), so what is verified here is the exchange itself."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from disco_b200.dist import all_gather_nodes, shard_range, tango_node_sharded


def test_shard_range_is_a_partition():
    for n in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle_step1(y, mask_z, **kw):
    """Stand-in for tango_step1 on CPU: float64 oracle, frame-major tensors like the GPU op."""
    from oracle import tango_f64
    B, _, C, L = y.shape
    Ys, zs = [], []
    for b in range(B):
        Y = np.array([tango_f64.stft64(y[b, 0, c].numpy()) for c in range(C)])          # (C, F, T)
        Rss, Rnn = tango_f64.masked_scm(Y, mask_z[b, 0].numpy().T)
        w = tango_f64.solve(Rss, Rnn)
        Ys.append(Y.transpose(0, 2, 1))
        zs.append(tango_f64.filter_sum(w, Y).T)
    return {"Y": torch.from_numpy(np.array(Ys))[:, None], "z_y": torch.from_numpy(np.array(zs))[:, None]}


def _oracle_step2(Y, Z, mask_w, nodes, **kw):
    """Z node-major [K, B, T, F] (what the all-gather delivers); one local node per rank here."""
    from oracle import tango_f64
    K, B = Z.shape[:2]
    node = list(nodes)[0]
    out = []
    for b in range(B):
        others = [j for j in range(K) if j != node]
        X = np.concatenate([Y[b, 0].numpy().transpose(0, 2, 1), Z[others, b].numpy().transpose(0, 2, 1)], axis=0)
        Rss, Rnn = tango_f64.masked_scm(X, mask_w[b, 0].numpy().T)
        out.append(tango_f64.filter_sum(tango_f64.solve(Rss, Rnn), X).T)
    return torch.from_numpy(np.array(out))[:, None]


def _worker(rank, world, port, L, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from disco_b200.synth import make_batch
        from oracle import tango_f64
        B, C = 3, 2
        y, s, n = make_batch(B, world, C, L, seed0=11)
        T, F = 1 + L // 256, 257
        masks = np.stack([[tango_f64.irm(tango_f64.stft64(s[b, k, 0]), tango_f64.stft64(n[b, k, 0])).T
                           for k in range(world)] for b in range(B)])                     # [B, K, T, F]
        # order check of the gather itself
        tag = torch.full((B, 3), float(rank))
        G = all_gather_nodes(tag)                       # one node per rank, [B, ...] form -> node-major [K, B, ...]
        assert G.shape == (world, B, 3) and all(float(G[k, 0, 0]) == k for k in range(world))
        zc = torch.full((B, 2, 2, 2), complex(rank, -rank), dtype=torch.complex128)     # [B, Kl = 2, T, F]
        zc[:, 1] += 10
        Gc = all_gather_nodes(zc)                       # two nodes per rank: node k = rank k // 2, local k % 2
        assert Gc.dtype == torch.complex128 and Gc.shape == (2 * world, B, 2, 2)
        assert Gc[2, 1, 0, 0] == complex(1, -1) and Gc[3, 0, 1, 1] == complex(11, -1) and Gc[1, 2, 0, 0] == complex(10, 0)
        ok = True
        calls = {"s1": [], "s2": []}

        def s1(y_, m_, ref_mic=0):                      # explicit signatures, like the GPU callables
            calls["s1"].append({"ref_mic": ref_mic})
            return _oracle_step1(y_, m_)

        def s2(Y_, Z_, m_, nodes, out_layout="TF"):
            calls["s2"].append({"out_layout": out_layout})
            return _oracle_step2(Y_, Z_, m_, nodes)
        for chunks in (1, 2):
            res = tango_node_sharded(torch.from_numpy(y[:, rank:rank + 1]), torch.from_numpy(masks[:, rank:rank + 1]),
                                     step1=s1, step2=s2, chunks=chunks, ref_mic=0, out_layout="TF")
            assert len(res["Z"]) == chunks and res["yf"].shape[0] == B
            Zall = torch.cat(res["Z"], dim=1)            # [K, B, T, F]
            for b in range(B):
                ref = tango_f64.offline_tango(y[b], s[b], n[b])
                err = np.linalg.norm(res["yf"][b, 0].numpy().T - ref["yf"][rank]) / np.linalg.norm(ref["yf"][rank])
                ok = ok and err < 1e-9
                errz = np.abs(Zall[:, b].numpy().transpose(0, 2, 1) - ref["z_y"]).max()
                ok = ok and errz < 1e-9
        # step-specific keywords reach only the step that takes them (ref_mic: step 1, out_layout: step 2);
        # passing both used to raise a TypeError in the other step
        ok = ok and len(calls["s1"]) == 3 and len(calls["s2"]) == 3
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_node_sharded_exchange_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 6000, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert got == [(0, True), (1, True)]
