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


class CrnnTangoPipeline:
    """The literal deployment chain of BASELINE configs[1] for single-node arrays, host to host: int16 PCM in pinned
    host memory -> device (half the bytes of float32; the conversion x / 32768 is exactly what soundfile's
    dtype='float32' read does, reference tango.py:95-100) -> |STFT| of the reference microphone -> the reference's
    CRNN mask estimators on the device (dnn_mask.estimate_masks_batch; tango.py:209-215) -> two-mask fused
    STFT+SCM, both solves, one-pass dual filter (captured CUDA graph) -> beamformed STFT back to the host.
    Only the signals cross PCIe on the way in.  The batch is cut into `chunks` slices on separate streams so
    upload, mask estimation, beamforming and download of different slices overlap."""

    def __init__(self, B, C, L, n_fft=512, chunks=4, device=None, models=None, exact=False, seed=0, bf16=False,
                 cudnn_benchmark=False):
        from . import dnn_mask
        if cudnn_benchmark:                        # static shapes: let cuDNN pick its convolution algorithms once
            torch.backends.cudnn.benchmark = True  # (process-wide switch, hence opt-in)
        self.bf16 = bool(bf16) and not exact
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.B, self.C, self.L, self.n_fft, self.exact = B, C, L, n_fft, exact
        chunks = max(1, min(chunks, B))
        self.splits = [((B * i) // chunks, (B * (i + 1)) // chunks) for i in range(chunks)]
        if models is None:           # randomly initialised networks of the reference architecture (tango.py:127-132)
            torch.manual_seed(seed)
            models = (dnn_mask.CRNN(1), dnn_mask.CRNN(1))
        self.models = tuple(m.to(self.device).eval() for m in models)
        self.plans = [TangoGraph(hi - lo, 1, C, L, n_fft=n_fft, chunks=1, device=self.device) for lo, hi in self.splits]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.splits]
        self.pcm = [torch.empty((hi - lo, 1, C, L), dtype=torch.int16, device=self.device) for lo, hi in self.splits]
        self.ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in self.splits]
        self.how = ("CrnnTangoPipeline: %d batch slices; per slice int16 H2D -> x/32768 -> STFT(ref mic) -> 2 CRNNs (%s) -> "
                    "graph replay -> D2H" % (chunks, "IEEE fp32" if exact else ("bf16 autocast" if self.bf16 else
                                                                                 "TF32 convolutions / matmuls")))

    @staticmethod
    def to_pcm(y_host):
        """float32 [-1, 1) signals -> int16 PCM (pinned), what a 16-bit wav file holds."""
        return (y_host.clamp(-1.0, 32767.0 / 32768.0) * 32768.0).round().to(torch.int16).pin_memory()

    def process(self, pcm_host, yf_host):
        """pcm_host [B, 1, C, L] int16 pinned, yf_host [B, 1, T, F] complex64 pinned."""
        from . import dnn_mask
        cur = torch.cuda.current_stream(self.device)
        for i, ((lo, hi), plan, st) in enumerate(zip(self.splits, self.plans, self.streams)):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                e = self.ev[i]
                e[0].record(st)
                self.pcm[i].copy_(pcm_host[lo:hi], non_blocking=True)
                torch.mul(self.pcm[i], 1.0 / 32768.0, out=plan.y)          # exact: a power of two
                e[1].record(st)
                Yref = ops.stft(plan.y[:, 0, 0].contiguous(), self.n_fft)  # [Bc, T, F]
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.bf16):
                    plan.mask_z[:, 0] = dnn_mask.estimate_masks_batch(self.models[0], Yref, exact=self.exact)
                    plan.mask_w[:, 0] = dnn_mask.estimate_masks_batch(self.models[1], Yref, exact=self.exact)
                e[2].record(st)
                plan.run()
                e[3].record(st)
                plan.store("yf", yf_host[lo:hi])
        for st in self.streams:
            cur.wait_stream(st)

    def report(self):
        """Mean per-slice times of the last process() call (call after a device synchronisation)."""
        up = [e[0].elapsed_time(e[1]) for e in self.ev]
        nn = [e[1].elapsed_time(e[2]) for e in self.ev]
        bf = [e[2].elapsed_time(e[3]) for e in self.ev]
        return {"slices": len(self.ev), "upload_convert_ms_per_slice": sum(up) / len(up),
                "crnn_ms_per_slice": sum(nn) / len(nn), "beamform_ms_per_slice": sum(bf) / len(bf),
                "crnn_ms_per_batch": sum(nn), "precision": "ieee fp32" if self.exact else ("bf16 autocast" if self.bf16 else "tf32"),
                "models": "2 x reference CRNN (517,729 parameters each), random weights"}
