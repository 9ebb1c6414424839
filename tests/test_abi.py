"""CPU-side checks of the drop-in boundary: the shared library builds, loads and exports every
symbol include/disco_b200.h declares; argument validation happens before any CUDA work."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from disco_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    from disco_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "disco_b200.h")).read()
    declared = set(re.findall(r"DISCO_API\s+[\w\s\*]+?\b(disco_\w+)\s*\(", hdr))
    assert len(declared) >= 15
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name
    assert lib.disco_abi_version() == 2


def test_n_frames_matches_reference_formula(lib):
    import numpy as np
    for L in (4000, 64000, 160000, 5001):
        for n_fft in (256, 512, 1024):
            hop = n_fft // 2
            assert lib.disco_n_frames(L, n_fft) == 3 + int(np.floor((L - n_fft) / hop))   # tango.py:287


def test_argument_validation_without_gpu(lib):
    from disco_b200 import _lib
    assert lib.disco_stft(None, None, 1, 4000, 500, None) == -1            # bad n_fft
    assert b"n_fft" in lib.disco_last_error()
    assert lib.disco_mwf_solve(None, None, None, None, 1, 4, 7, 1, 1.0, None) == -1
    assert b"Unknown filter reference" in lib.disco_last_error()
    assert lib.disco_mwf_solve(None, None, None, None, 1, 17, 0, 1, 1.0, None) == -2
    assert lib.disco_tf_mask(None, None, None, 10, 5, 1, 0.0, None) == -1
    assert lib.disco_stft_scm_workspace(64, 4, 160000, 512) > 0
    assert lib.disco_stft_scm2_workspace(64, 4, 160000, 512) == 2 * lib.disco_stft_scm_workspace(64, 4, 160000, 512)
    # which (n_fft, channels, masks) the fused STFT+SCM kernel covers: up to 8 mics, two masks up to 4 mics
    sup = lib.disco_stft_scm_supported
    assert sup(512, 8, 1) and sup(256, 5, 1) and sup(512, 4, 2) and sup(1024, 4, 1)
    assert not sup(1024, 8, 1) and not sup(512, 8, 2) and not sup(512, 9, 1) and not sup(1024, 4, 2)
    assert lib.disco_filter_dual(None, None, None, None, None, None, 0, 0, 1, 4, 10, 512, None) == -1
    assert lib.disco_masked_scm(None, None, None, 0, None, None, 1, 2, 2, 10, 512, None, 0, 7, None) == -1   # bad z_layout
    assert b"z_layout" in lib.disco_last_error()
    with pytest.raises(_lib.DiscoError):
        _lib.check(lib.disco_init(300))


def test_no_oracle_import_in_product():
    """The product package must never import the oracle (or route through any CPU fallback)."""
    pkg = os.path.join(ROOT, "disco_b200")
    pat = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b|from\s+\.+oracle\b)", re.M)
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert not pat.search(src), fn
                assert "/root/reference" not in src, fn
