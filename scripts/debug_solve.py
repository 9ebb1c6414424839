import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from disco_b200 import ops
from oracle import tango_f64
dev = torch.device("cuda:0")
for D in (12, 13, 14, 15, 16):
    rng = np.random.default_rng(D)
    n = 300
    def hpd(r):
        a = rng.standard_normal((n, D, r)) + 1j * rng.standard_normal((n, D, r))
        return a @ a.conj().transpose(0, 2, 1) / r
    Rss = (hpd(D + 2) * 0.1 + 3 * hpd(1)).astype(np.complex64)
    Rnn = hpd(D + 3).astype(np.complex64)
    W, t1 = ops.mwf_solve(torch.from_numpy(Rss).to(dev), torch.from_numpy(Rnn).to(dev), 1.0, "gevd", 1)
    wref, tref, lam = tango_f64.gevd_filter(Rss.astype(np.complex128), Rnn.astype(np.complex128), 1.0, 1)
    W, t1 = W.cpu().numpy(), t1.cpu().numpy()
    ew = np.linalg.norm(W - wref, axis=1) / np.linalg.norm(wref, axis=1)
    et = np.linalg.norm(t1 - tref, axis=1) / np.linalg.norm(tref, axis=1)
    print(D, "W err max/median", ew.max(), np.median(ew), "t1 err", et.max(), np.median(et), "bad", (et > 1e-4).sum(),
          "lam ratio", np.median(lam[:, 0] / lam[:, 1]))
    bad = np.where(et > 1e-4)[0][:3]
    for b in bad:
        print("   idx", b, "lam", lam[b, :4], "ratio t1/tref", (t1[b] / tref[b])[:3])
