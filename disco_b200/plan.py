"""CUDA-graph execution plan of the batched Tango path.

The path is a handful of launches per batch, several of them tens of microseconds long, so launch
gaps and the latency-bound per-bin solves matter.  ``TangoGraph`` captures the whole two-step
pipeline for a fixed problem shape once and replays it:

* static device buffers for the inputs (signals, masks) and outputs -- `load()` copies new data in
  (from pinned host memory or device tensors, asynchronously), `run()` replays;
* the batch is split into `chunks` independent utterance groups captured on parallel stream
  branches, so the float64 per-bin solve of one chunk (few warps, latency-bound) overlaps the
  HBM-bound streaming kernels of another.

Deployment mode only (masks supplied, e.g. by a DNN): y [B, K, C, L], masks [B, K, T, F] frame-major.
"""
import torch

from . import ops
from .tango import tango_batched


class TangoGraph:
    def __init__(self, B, K, C, L, n_fft=512, chunks=2, device=None, shared_mask=False, **tango_kw):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.shape = (B, K, C, L)
        self.n_fft = n_fft
        T, F = ops.n_frames(L, n_fft), n_fft // 2 + 1
        self.T, self.F = T, F
        chunks = max(1, min(chunks, B))
        self.splits = [((B * i) // chunks, (B * (i + 1)) // chunks) for i in range(chunks)]
        self.y = torch.zeros((B, K, C, L), dtype=torch.float32, device=self.device)
        self.mask_z = torch.full((B, K, T, F), 0.5, dtype=torch.float32, device=self.device)
        self.mask_w = self.mask_z if shared_mask else torch.full_like(self.mask_z, 0.5)
        kw = dict(n_fft=n_fft, out_layout="TF", diagnostics=False)
        kw.update(tango_kw)
        self._kw = kw
        ops.init(n_fft)
        self.outputs = None
        self.graph = None
        self._capture()

    def _chunk(self, lo, hi):
        return tango_batched(self.y[lo:hi], masks=(self.mask_z[lo:hi], self.mask_w[lo:hi]), **self._kw)

    def _capture(self):
        warm = torch.cuda.Stream(device=self.device)
        warm.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(warm):                      # eager warm-up: tables, function attributes
            for lo, hi in self.splits:
                self._chunk(lo, hi)
        torch.cuda.current_stream(self.device).wait_stream(warm)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        side = [torch.cuda.Stream(device=self.device) for _ in self.splits[1:]]
        outs = [None] * len(self.splits)
        with torch.cuda.graph(g):
            main = torch.cuda.current_stream(self.device)
            for st in side:                                  # fork
                st.wait_stream(main)
            for i, (lo, hi) in enumerate(self.splits):
                if i == 0:
                    outs[0] = self._chunk(lo, hi)
                else:
                    with torch.cuda.stream(side[i - 1]):
                        outs[i] = self._chunk(lo, hi)
            for st in side:                                  # join
                main.wait_stream(st)
        self.graph = g
        self.outputs = outs

    def load(self, y=None, mask_z=None, mask_w=None):
        """Copy new inputs into the static buffers (async on the current stream; pinned host or device)."""
        if y is not None:
            self.y.copy_(y, non_blocking=True)
        if mask_z is not None:
            self.mask_z.copy_(mask_z, non_blocking=True)
        if mask_w is not None and self.mask_w is not self.mask_z:
            self.mask_w.copy_(mask_w, non_blocking=True)

    def run(self):
        """Replay the captured pipeline on the current stream; returns the per-chunk output dicts
        (static tensors, valid until the next run)."""
        self.graph.replay()
        return self.outputs

    def output(self, name):
        """Concatenate one output over the chunks (allocates; use `outputs` to avoid the copy)."""
        return torch.cat([o[name] for o in self.outputs], dim=0)

    def store(self, name, host_out):
        """Copy one output into a [B, ...] (pinned) host tensor, chunk by chunk, asynchronously."""
        for (lo, hi), o in zip(self.splits, self.outputs):
            host_out[lo:hi].copy_(o[name], non_blocking=True)


class TangoPipeline:
    """Host-to-host execution: pinned host signals / masks in, beamformed STFT out, with the PCIe
    copies overlapped.  The batch is cut into `chunks` slices, each with its own captured TangoGraph
    and its own stream; slice i's host->device copy, compute and device->host copy are enqueued
    back to back on stream i, so while one slice computes, the next one is uploading and the
    previous one is downloading (both PCIe directions busy).  The path is PCIe-bound end to end
    (328 MB per 64 x 4-mic x 10 s batch vs 0.37 ms of GPU time), so this is what sets `e2e`."""

    def __init__(self, B, K, C, L, n_fft=512, chunks=4, device=None, **tango_kw):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        chunks = max(1, min(chunks, B))
        self.splits = [((B * i) // chunks, (B * (i + 1)) // chunks) for i in range(chunks)]
        self.plans = [TangoGraph(hi - lo, K, C, L, n_fft=n_fft, chunks=1, device=self.device, **tango_kw)
                      for lo, hi in self.splits]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.splits]

    def process(self, y_host, mask_z_host, mask_w_host, yf_host):
        """Enqueue one batch; returns after enqueueing (synchronise the device or the output's consumer
        stream before reading yf_host).  All host tensors should be pinned."""
        cur = torch.cuda.current_stream(self.device)
        for (lo, hi), plan, st in zip(self.splits, self.plans, self.streams):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                plan.load(y_host[lo:hi], mask_z_host[lo:hi], None if mask_w_host is None else mask_w_host[lo:hi])
                plan.run()
                plan.store("yf", yf_host[lo:hi])
        for st in self.streams:
            cur.wait_stream(st)
