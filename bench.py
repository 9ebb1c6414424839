#!/usr/bin/env python
"""Benchmark of the B200 MWF beamforming hot path (BASELINE.json metric: beamformed frames/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|...]

One "step" = one pass of the whole two-step Tango path (STFT -> masked SCM -> per-bin GEVD-MWF
-> filter-and-sum, twice) over one batch of synthetic utterances.  Default workload = BASELINE
configs[1]: 1 node x 4 mics, batch = 64 x 10 s @ 16 kHz per GPU, 512-pt STFT, device-resident
DNN-style masks.  Frame unit: one STFT frame of one node's beamformed output, so a
step produces B*K*T frames per GPU.  N > 1: utterances shard over ranks, no data-path collective
(weak scaling); timing = max over ranks of CUDA-event time between barriers.

The JSON line also carries
  roofline      the fused stft_scm kernel: algorithmic bytes (SURVEY.md 8d:
                4CL + 4FT + 8CFT + 16FC^2 per group) / its CUDA-event time, vs the measured HBM peak
  e2e           the same metric through the public API with pinned HOST buffers (H2D of signals and
                masks, D2H of the beamformed STFT inside the timed region)
  cpu_baseline  the oracle port of the reference (oracle/tango_np.py, per-frame np.outer loops
                like tango.py:357-374) on the host cores, bounded sample
--impl reference times only that CPU path (the reference is pure Python/NumPy and cannot travel to
the GPU box; its restatement is pinned to the reference's outputs by tests/test_oracle.py).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B per GPU, K nodes, C mics, L samples, n_fft, description)
    "cfg1": (1, 1, 2, 64000, 512, "1 node x 2 mics, 4 s, 512-pt STFT (BASELINE configs[0])"),
    "cfg2": (64, 1, 4, 160000, 512, "1 node x 4 mics, batch=64 x 10 s per GPU, 512-pt STFT, DNN mask (BASELINE configs[1])"),
    "cfg3": (64, 4, 4, 160000, 512, "4 nodes x 4 mics, batch=64 x 10 s per GPU, all nodes on-GPU (BASELINE configs[2] shape)"),
    "cfg5": (64, 8, 2, 160000, 512, "8 nodes x 2 mics, batch=64 x 10 s per GPU (BASELINE configs[4] shape)"),
    "cfg4_256": (128, 1, 8, 160000, 256, "8 mics, 256-pt STFT, batch=128 x 10 s per GPU (BASELINE configs[3] sweep point)"),
    "cfg4_512": (128, 1, 8, 160000, 512, "8 mics, 512-pt STFT, batch=128 x 10 s per GPU (BASELINE configs[3] sweep point)"),
    "cfg4_1024": (128, 1, 8, 160000, 1024, "8 mics, 1024-pt STFT, batch=128 x 10 s per GPU (BASELINE configs[3] sweep point)"),
}


def stft_scm_bytes(C, L, n_fft):
    F, T = n_fft // 2 + 1, 1 + L // (n_fft // 2)
    return 4 * C * L + 4 * F * T + 8 * C * F * T + 16 * F * C * C


# ----------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference), bounded sample, one utterance per process
# ----------------------------------------------------------------------------------------------
def _cpu_one(args):
    seed, K, C, L, n_fft, gran = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from disco_b200.synth import make_utterance
    from oracle import tango_np
    y, s, n = make_utterance(seed, K, C, L)
    t0 = time.perf_counter()
    tango_np.offline_tango(list(y), list(s), list(n), n_fft=n_fft, n_hop=n_fft // 2, granularity=gran)
    return time.perf_counter() - t0


def _cpu_warm(_):
    import scipy.linalg  # noqa: F401
    from oracle import tango_np  # noqa: F401
    time.sleep(0.2)
    return 0


def usable_cores():
    """Host cores this process may actually use: affinity mask and cgroup CPU quota, capped at 64
    workers so the bounded sample stays bounded."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, min(n, 64))


def cpu_baseline(K, C, L, n_fft, budget_s=20.0, granularity="frame", procs=None):
    """frames/s of the oracle port with one utterance per host process (how the reference
    parallelises: exp/ex1/loop_tango.sh launches one process per utterance)."""
    import multiprocessing as mp
    cores = procs or usable_cores()
    T = 1 + L // (n_fft // 2)
    # bound the sample: shorten the utterance so one of them takes <~ budget (the reference runs ~120 frames/s/core)
    est_rate = 110.0 if granularity == "frame" else 20000.0
    max_frames = max(64, int(budget_s * est_rate / K))
    Ls = min(L, (max_frames - 1) * (n_fft // 2))
    Ts = 1 + Ls // (n_fft // 2)
    ctx = mp.get_context("spawn")
    os.environ["OMP_NUM_THREADS"] = "1"          # inherited by the workers: one thread per process
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_warm, range(cores), chunksize=1)          # interpreter + NumPy/SciPy imports: untimed
        t0 = time.perf_counter()
        pool.map(_cpu_one, [(1000 + i, K, C, Ls, n_fft, granularity) for i in range(cores)], chunksize=1)
        wall = time.perf_counter() - t0
    return {"value": cores * K * Ts / wall, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": "%d utterances (one per process) x %d node(s) x %d mics x %.2f s (%d frames), oracle/tango_np.py "
                      "granularity=%s, workers pre-started" % (cores, K, C, Ls / 16000.0, Ts, granularity),
            "seconds": wall, "frames": cores * K * Ts}


# ----------------------------------------------------------------------------------------------
def clock_sampler(stop, out, gpu_index):
    """Sample SM clock and throttle reasons through NVML every ~2 ms while the timed region runs."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
        h = nv.nvmlDeviceGetHandleByIndex(phys)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not stop.is_set():
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            out.append([str(sm), str(mx)] + ["Active" if r & bits[k] else "Not Active" for k in
                                             ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
            stop.wait(0.002)
        return
    except Exception:
        pass
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5)
            parts = [p.strip() for p in r.stdout.strip().split(",")]
            if len(parts) >= 6:
                out.append(parts)
        except Exception:
            pass
        stop.wait(0.1)


def summarize_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = sorted(float(s[0]) for s in samples)
    reasons = []
    for i, nm in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
        if any(s[2 + i].lower().startswith("active") for s in samples):
            reasons.append(nm)
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(samples[0][1]), "reasons": reasons, "samples": len(sm)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override utterances per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="batch slices of the host-to-host pipeline")
    ap.add_argument("--chunks", type=int, default=1, help="batch chunks captured on parallel graph branches")
    args = ap.parse_args()
    B, K, C, L, n_fft, desc = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    T, F = 1 + L // (n_fft // 2), n_fft // 2 + 1
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "%s: %s" % (args.workload, desc), "nodes": K, "mics_per_node": C, "utterance_s": L / 16000.0,
              "n_fft": n_fft, "hop": n_fft // 2, "batch_per_gpu": B, "global_batch": B * world, "frames_per_step": B * K * T * world,
              "mask": "device-resident synthetic DNN-style masks U[0,1], frame-major (T,F)", "parallelism": "utterance-sharded x%d, no collective" % world, "execution": "CUDA graph, %d batch chunk(s) on parallel branches" % max(1, min(args.chunks, B)),
              "l2": "inputs larger than L2 (y %.0f MB, Y %.0f MB per GPU)" % (B * K * C * L * 4 / 1e6, B * K * C * T * F * 8 / 1e6)}

    if args.impl == "reference":
        if rank != 0:
            return
        t_all, frames = 0.0, 0
        last = None
        for i in range(max(1, min(args.steps, 3))):          # each step = one bounded sample on all host cores
            last = cpu_baseline(K, C, L, n_fft, budget_s=15.0)
            t_all += last["seconds"]
            frames += last["frames"]
        val = frames / t_all
        last["value"] = val
        print(json.dumps({"impl": "reference", "metric": "beamformed frames/sec (16kHz, 512-pt STFT)", "value": val,
                          "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * t_all / max(1, min(args.steps, 3)), "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "c64 SCM/cggev, c128 filters (reference dtype flow)", "data": "synthetic",
                          "config": config, "cpu_baseline": last,
                          "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    from disco_b200 import ops
    from disco_b200.tango import tango_batched
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic inputs (seeded): a few distinct utterances tiled to the batch, masks U[0,1]
    from disco_b200.synth import make_utterance
    base = [make_utterance(1000 * rank + i, K, C, L)[0] for i in range(4)]
    y_host = torch.from_numpy(np.stack([base[i % 4] for i in range(B)])).pin_memory()      # [B,K,C,L]
    g = torch.Generator().manual_seed(1234 + rank)
    mz_host = torch.rand((B, K, T, F), generator=g, dtype=torch.float32).pin_memory()
    mw_host = torch.rand((B, K, T, F), generator=g, dtype=torch.float32).pin_memory()
    y, mz, mw = y_host.to(dev), mz_host.to(dev), mw_host.to(dev)
    ops.init(n_fft)
    from disco_b200.plan import TangoGraph
    chunks = max(1, min(args.chunks, B))
    # stft_scm, mwf_solve, fused middle pass (or filter_sum + masked_scm), mwf_solve, filter_sum
    kernels_per_chunk = 5 if (K == 1 or ops.tango_mid_supported(C, K)) else 6
    launches_per_step = kernels_per_chunk * chunks
    plan = TangoGraph(B, K, C, L, n_fft=n_fft, chunks=chunks, device=dev)   # CUDA graph of the whole step
    plan.load(y, mz, mw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region 1: whole-path throughput, inputs resident in HBM (graph replays)
    for _ in range(args.warmup):
        plan.run()
    barrier()
    samples, stop = [], threading.Event()
    th = threading.Thread(target=clock_sampler, args=(stop, samples, local_rank), daemon=True)
    if rank == 0:
        th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        plan.run()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    tmax = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    frames = B * K * T * world * args.steps
    value = frames / (ms / 1e3)

    # ---- timed region 2: the fused stft_scm op alone, CUDA events on its launch stream inside eager steps
    import disco_b200.tango as tango_mod
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    cur = {"i": -1}
    orig = ops.stft_scm

    def timed_stft_scm(*a, **k):
        i = cur["i"]
        if i < 0:
            return orig(*a, **k)
        ev[i][0].record()
        r = orig(*a, **k)
        ev[i][1].record()
        return r
    tango_mod.ops.stft_scm = timed_stft_scm

    def eager_step():
        return tango_batched(y, masks=(mz, mw), n_fft=n_fft, out_layout="TF", diagnostics=False)
    for _ in range(3):
        eager_step()
    barrier()
    n_k = min(args.steps, 50)
    for i in range(n_k):
        cur["i"] = i
        eager_step()
    barrier()
    cur["i"] = -1
    tango_mod.ops.stft_scm = orig
    fused = C <= 4          # larger nodes run stft + masked_scm (the fused kernel holds C <= 4)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in ev[:n_k]])) if fused else None

    # ---- e2e: pinned host buffers in, beamformed STFT out, through the public API (plan.load/run/store)
    e2e = None
    if not args.no_e2e:
        yf_host = torch.empty((B, K, T, F), dtype=torch.complex64).pin_memory()

        from disco_b200.plan import TangoPipeline
        pipe = TangoPipeline(B, K, C, L, n_fft=n_fft, chunks=args.e2e_chunks, device=dev)

        def e2e_step():       # H2D of signals + masks, the whole path, D2H of yf -- overlapped across batch slices
            pipe.process(y_host, mz_host, mw_host, yf_host)
        for _ in range(2):
            e2e_step()
        barrier()
        n_e2e = max(3, min(args.steps // 4, 50))
        e0.record()
        for _ in range(n_e2e):
            e2e_step()
        e1.record()
        barrier()
        t2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e = {"value": B * K * T * world * n_e2e / (float(t2.item()) / 1e3), "unit": "frames/s",
               "h2d_bytes_per_step": int(y_host.numel() * 4 + 2 * mz_host.numel() * 4),
               "d2h_bytes_per_step": int(yf_host.numel() * 8), "steps": n_e2e,
               "how": "TangoPipeline: %d batch slices, per-slice H2D -> graph replay -> D2H on its own stream (pinned host buffers)" % args.e2e_chunks}
    stop.set()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        alg = stft_scm_bytes(C, L, n_fft) * B * K
        traffic = None       # dram__bytes_read.sum + dram__bytes_write.sum of the kernel, from the committed ncu capture
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if tj.get("workload") == args.workload and tj.get("batch") == B:
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        if kern_ms is None:
            roof = {"bound": "hbm", "kernel": None, "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
                    "note": "C > 4: stft and masked_scm run as two kernels; no fused-kernel roofline for this workload"}
        else:
          achieved = alg / (kern_ms / 1e3) / 1e9
          roof = {"bound": "hbm", "kernel": "stft_scm_kernel<%d,%d,true>" % (n_fft, C), "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                "algorithmic_bytes_per_launch": alg, "kernel_ms": kern_ms, "share_of_step": kern_ms / (ms / args.steps),
                "timed": "CUDA events around the op in %d eager steps (the throughput region replays a CUDA graph)" % n_k}
        res = {"metric": "beamformed frames/sec (16kHz, 512-pt STFT)", "value": value, "unit": "frames/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32 (c64 spectra, f32 SCM accumulation, f64 per-bin solve)",
               "data": "synthetic", "config": config, "clocks": summarize_clocks(samples), "gpu_launches": launches_per_step * args.steps,
               "roofline": roof}
        if e2e:
            res["e2e"] = e2e
        if not args.no_cpu and world == 1:      # the CPU leg is an N = 1 measurement (rank 0 owns the whole host)
            res["cpu_baseline"] = cpu_baseline(K, C, L, n_fft, budget_s=12.0)
            res["cpu_baseline_vectorized"] = cpu_baseline(K, C, L, n_fft, budget_s=6.0, granularity="bin")
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
