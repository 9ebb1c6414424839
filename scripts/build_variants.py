"""Build alternative tunings of single translation units into build/variants/libdisco_b200_<name>.so
(every other object is reused from the regular build) for A/B timing on the GPU box:

    python scripts/build_variants.py            # builds every variant listed in VARIANTS
    DISCO_B200_LIB=build/variants/libdisco_b200_<name>.so python scripts/quick_check.py

The shipped library is always disco_b200/libdisco_b200.so built without any -D switch."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from disco_b200 import build as B  # noqa: E402

OUT = os.path.join(ROOT, "build", "variants")
# name: (source file, [-D flags])
VARIANTS = {
    "fd_uf4": ("filter_dual.cu", ["-DDISCO_FD_UF=4"]),
    "fd_uf4_b3": ("filter_dual.cu", ["-DDISCO_FD_UF=4", "-DDISCO_FD_MINB=3"]),
    "fd_stcs": ("filter_dual.cu", ["-DDISCO_FD_STCS=1"]),
    "fd_uf4_stcs": ("filter_dual.cu", ["-DDISCO_FD_UF=4", "-DDISCO_FD_STCS=1"]),
    "fd_b3": ("filter_dual.cu", ["-DDISCO_FD_MINB=3"]),
    "fd_want32": ("filter_dual.cu", ["-DDISCO_FD_WANT=32"]),
    "fd_want32_uf4": ("filter_dual.cu", ["-DDISCO_FD_WANT=32", "-DDISCO_FD_UF=4"]),
    "fd_want16_stcs": ("filter_dual.cu", ["-DDISCO_FD_WANT=16", "-DDISCO_FD_STCS=1"]),
    "ss_pf2": ("stft_scm.cu", ["-DDISCO_SS_PF=2"]),
    "ss_pf4": ("stft_scm.cu", ["-DDISCO_SS_PF=4"]),
    "ss_fw4": ("stft_scm.cu", ["-DDISCO_SS_FW=4"]),
    "ss_fg2": ("stft_scm.cu", ["-DDISCO_SS_FG=2"]),
    "ss_ffthi": ("stft_scm.cu", ["-DDISCO_SS_FFTHI=1"]),
    "ss_fg2_ffthi": ("stft_scm.cu", ["-DDISCO_SS_FG=2", "-DDISCO_SS_FFTHI=1"]),
    "ss_fg2_pf2": ("stft_scm.cu", ["-DDISCO_SS_FG=2", "-DDISCO_SS_PF=2"]),
    # written at the end of round 2, not yet run on a GPU: squaring entries dealt to all lanes of a solver group
    "sv_spread": ("solve.cu", ["-DDISCO_SOLVE_SPREAD=1"]),
}


def main(names=None):
    B.build()
    os.makedirs(OUT, exist_ok=True)
    for name, (src, flags) in VARIANTS.items():
        if names and name not in names:
            continue
        obj = os.path.join(OUT, name + ".o")
        cmd = [B.NVCC] + B.FLAGS + flags + ["-c", os.path.join(B.CSRC, src), "-o", obj]
        subprocess.run(cmd, check=True)
        objs = [obj if s == src else os.path.join(B.OBJ, s.replace(".cu", ".o")) for s in B.SOURCES]
        lib = os.path.join(OUT, "libdisco_b200_%s.so" % name)
        subprocess.run([B.NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"], check=True)
        print(lib)


if __name__ == "__main__":
    main(set(sys.argv[1:]) or None)
