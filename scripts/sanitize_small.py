"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from disco_b200 import ops
from disco_b200.synth import make_batch
from disco_b200.tango import tango_batched
dev = torch.device("cuda:0")
for (B, K, C, L, n_fft) in ((3, 2, 4, 9000, 512), (2, 1, 3, 5003, 256), (2, 3, 2, 6000, 1024)):
    y, s, n = make_batch(B, K, C, L, seed0=1)
    out = tango_batched(torch.from_numpy(y).to(dev), torch.from_numpy(s).to(dev), torch.from_numpy(n).to(dev), n_fft=n_fft)
    x = ops.istft(ops.stft(torch.from_numpy(y).to(dev), n_fft), L, n_fft)
    torch.cuda.synchronize()
    print("ok", B, K, C, L, n_fft, float(out["yf"].abs().mean()), float((x.cpu() - torch.from_numpy(y)).abs().max()))
