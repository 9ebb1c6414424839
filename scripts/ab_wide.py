"""Per-op CUDA-event timing of the wide-channel kernels on BASELINE shapes (AB_ONLY=<name> restricts the shapes).
Prints microseconds per launch (median of 20 after 3 warm-up launches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from disco_b200 import ops  # noqa: E402


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
    return 1e3 * ts[n // 2]


def main():
    dev = torch.device("cuda:0")
    T, F = 626, 257
    gen = torch.Generator(device=dev).manual_seed(1)
    cplx = lambda *s: torch.view_as_complex(torch.randn(*s, 2, device=dev, generator=gen))
    tag = "r1"
    only = os.environ.get("AB_ONLY")
    for name, B, K, C in [("cfg3", 64, 4, 4), ("cfg5", 64, 8, 2), ("cfg4", 128, 1, 8), ("k2c4", 64, 2, 4)]:
        if only and name != only:
            continue
        Y, W = cplx(B, K, C, T, F), cplx(B, K, F, C)
        m = torch.rand(B, K, T, F, device=dev, generator=gen)
        if K > 1:
            us = timeit(lambda: ops.tango_mid(W, Y, m, ref=0))
            print(f"{tag} {name} tango_mid B={B} K={K} C={C}: {us:8.1f} us", flush=True)
            z = ops.tango_mid(W, Y, m, ref=0)[0]
            us = timeit(lambda: ops.masked_scm(Y, m, z))
            print(f"{tag} {name} masked_scm D={C + K - 1} (unfused step 2): {us:8.1f} us", flush=True)
        else:
            us = timeit(lambda: ops.filter_sum_scm(W, Y, m, ref=0))
            print(f"{tag} {name} filter_sum_scm D={C}: {us:8.1f} us", flush=True)
            us = timeit(lambda: ops.masked_scm(Y, m, None))
            print(f"{tag} {name} masked_scm D={C}: {us:8.1f} us", flush=True)
        del Y, W, m
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
