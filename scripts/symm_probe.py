"""Probe: does torch's symmetric memory (peer-mapped buffers, NVSwitch multicast) work on this box?

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/symm_probe.py

Prints, per rank: the peer pointers, the multicast pointer (0 = no NVLS), and whether a value written into a peer's
buffer arrives.  Diagnostic only -- the shipped exchange is the NCCL all-gather of disco_b200/dist.py."""
import os
import sys

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    try:
        import torch.distributed._symmetric_memory as symm_mem
    except Exception as exc:
        print("rank %d: no symmetric-memory module: %r" % (rank, exc))
        return 0
    try:
        t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
        hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
        print("rank %d: buffer_ptrs %s multicast_ptr %s signal_pads %s" % (
            rank, [hex(p) for p in hdl.buffer_ptrs], hex(getattr(hdl, "multicast_ptr", 0) or 0),
            [hex(p) for p in getattr(hdl, "signal_pad_ptrs", [])]))
        t.fill_(float(rank))
        hdl.barrier()
        peer = (rank + 1) % world
        pbuf = hdl.get_buffer(peer, (1 << 20,), torch.float32)
        pbuf[:1024].fill_(100.0 + rank)                      # a P2P store into the neighbour's buffer
        hdl.barrier()
        torch.cuda.synchronize()
        got = float(t[0].item())
        print("rank %d: own buffer now starts with %.1f (expected %.1f)" % (rank, got, 100.0 + (rank - 1) % world))
        # bandwidth of a plain peer copy (copy kernel over NVLink), 256 MB
        big = symm_mem.empty(64 << 20, dtype=torch.float32, device=dev)
        hb = symm_mem.rendezvous(big, dist.group.WORLD.group_name)
        src = torch.ones(64 << 20, dtype=torch.float32, device=dev)
        dstp = hb.get_buffer(peer, (64 << 20,), torch.float32)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            dstp.copy_(src)
        hb.barrier()
        e0.record()
        for _ in range(5):
            dstp.copy_(src)
        e1.record()
        torch.cuda.synchronize()
        print("rank %d: peer copy %.0f GB/s" % (rank, 5 * 256e6 * 1.048576 / (e0.elapsed_time(e1) * 1e-3) / 1e9))
    except Exception as exc:
        print("rank %d: symmetric memory failed: %r" % (rank, exc))
    dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
