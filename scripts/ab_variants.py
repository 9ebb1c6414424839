"""A/B timing of kernel tunings on the GPU box: every library in build/variants (scripts/build_variants.py)
plus the shipped one, each in its own process, CUDA-event timings of the kernels the variant touches.

    python scripts/ab_variants.py [prefix ...]        # e.g. fd_  or  ss_
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VAR = os.path.join(ROOT, "build", "variants")


def child(what):
    import torch
    sys.path.insert(0, ROOT)
    from disco_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)

    def timeit(fn, n=30, warm=5):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3

    out = {}
    L, n_fft = 160000, 512
    T, F = 1 + L // 256, 257
    cplx = lambda *s: torch.complex(torch.randn(s, generator=g), torch.randn(s, generator=g)).to(dev)
    if what.startswith("fd"):
        G, C = 64, 4
        Y, W1, W2 = cplx(G, 1, C, T, F), cplx(G, 1, F, C), cplx(G, 1, F, C)
        out["filter_dual<4> TF us"] = timeit(lambda: ops.filter_dual(W1, W2, Y))
        out["filter_dual<4> FT us"] = timeit(lambda: ops.filter_dual(W1, W2, Y, out_layout="FT"))
        Y2, V1, V2 = cplx(128, 1, 2, T, F), cplx(128, 1, F, 2), cplx(128, 1, F, 2)
        out["filter_dual<2> TF us"] = timeit(lambda: ops.filter_dual(V1, V2, Y2))
    else:
        for (G, C, nm) in ((64, 4, 2), (64, 4, 1), (128, 8, 1), (128, 2, 2)):
            x = torch.randn((G, C, L), generator=g).to(dev)
            m, m2 = torch.rand((G, T, F), generator=g).to(dev), torch.rand((G, T, F), generator=g).to(dev)
            if nm == 2:
                out["stft_scm<512,%d,2> G=%d us" % (C, G)] = timeit(lambda: ops.stft_scm2(x, m, m2))
            else:
                out["stft_scm<512,%d,1> G=%d us" % (C, G)] = timeit(lambda: ops.stft_scm(x, m, keep_partials=True))
            del x, m, m2
    print(json.dumps(out))


def main():
    prefixes = sys.argv[1:] or [""]
    libs = [("shipped", None)]
    if os.path.isdir(VAR):
        for fn in sorted(os.listdir(VAR)):
            if fn.endswith(".so"):
                name = fn[len("libdisco_b200_"):-3]
                if any(name.startswith(p) for p in prefixes):
                    libs.append((name, os.path.join(VAR, fn)))
    for name, path in libs:
        kinds = {"fd" if (name.startswith("fd") or (name == "shipped" and any(p.startswith("fd") for p in prefixes))) else "ss"}
        if name == "shipped" and prefixes == [""]:
            kinds = {"fd", "ss"}
        elif name == "shipped":
            kinds = {("fd" if p.startswith("fd") else "ss") for p in prefixes}
        for kind in sorted(kinds):
            env = dict(os.environ)
            if path:
                env["DISCO_B200_LIB"] = path
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", kind], env=env, capture_output=True, text=True,
                               timeout=600)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]
            print("%-18s %s" % (name, line), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
    else:
        main()
