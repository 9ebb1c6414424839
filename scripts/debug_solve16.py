import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from disco_b200 import ops
from oracle import tango_f64
dev = torch.device("cuda:0")
D = 16
rng = np.random.default_rng(D)
n = 64
def hpd(r):
    a = rng.standard_normal((n, D, r)) + 1j * rng.standard_normal((n, D, r))
    return a @ a.conj().transpose(0, 2, 1) / r
Rss = (hpd(D + 2) * 0.1 + 3 * hpd(1)).astype(np.complex64)
Rnn = hpd(D + 3).astype(np.complex64)
for typ in ("gevd", "mwf", "r1-mwf"):
    W, t1 = ops.mwf_solve(torch.from_numpy(Rss).to(dev), torch.from_numpy(Rnn).to(dev), 1.0, typ, 1)
    wref = tango_f64.solve(Rss.astype(np.complex128), Rnn.astype(np.complex128), 1.0, typ)
    W = W.cpu().numpy()
    ew = np.linalg.norm(W - wref, axis=1) / np.linalg.norm(wref, axis=1)
    print(os.environ.get("DISCO_SOLVE_ALT"), typ, "W err max/median", ew.max(), np.median(ew))
# identity matrices: Rnn = I -> plain eig of Rss
I = np.broadcast_to(np.eye(D, dtype=np.complex64), (n, D, D)).copy()
W, t1 = ops.mwf_solve(torch.from_numpy(Rss).to(dev), torch.from_numpy(I).to(dev), 1.0, "gevd", 1)
wref = tango_f64.solve(Rss.astype(np.complex128), I.astype(np.complex128), 1.0, "gevd")
print("Rnn=I gevd err", (np.linalg.norm(W.cpu().numpy() - wref, axis=1) / np.linalg.norm(wref, axis=1)).max())
W, t1 = ops.mwf_solve(torch.from_numpy(I).to(dev), torch.from_numpy(Rnn).to(dev), 1.0, "gevd", "full")
wref = tango_f64.gevd_filter(I.astype(np.complex128), Rnn.astype(np.complex128), 1.0, "full")[0]
print("Rss=I gevd full err", (np.linalg.norm(W.cpu().numpy() - wref, axis=1) / np.linalg.norm(wref, axis=1)).max())
