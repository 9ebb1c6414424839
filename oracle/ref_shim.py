"""Import the UNMODIFIED reference hot path from /root/reference — TEST INFRASTRUCTURE.

Only usable where the reference is mounted (the authoring container).  Nothing is
copied: the reference modules are imported from where they lie, after stubbing the
third-party packages they import but which are absent here:

  soundfile, ipdb, matplotlib(.pyplot/.patches), acoustics.signal.OctaveBand,
  mir_eval.separation.bss_eval_sources, pystoi.stoi.stoi          -> empty stubs
  disco_theque.dnn.models.heymann.build_heymann (missing in the repo, tango.py:20)
  librosa.core.stft / istft                                       -> oracle.librosa_np
  dnn.utils <-> dnn.models.crnn circular import                   -> pre-seeded stub
  dnn.utils.move_to_device (imported at speech_enhancement/utils.py:5, defined nowhere)
  np.int (removed in NumPy >= 1.24; sigproc_utils.py:53)          -> int

``load()`` returns the namespace of reference callables used to pin the oracle.
"""
import importlib
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("DISCO_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "disco_theque"))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


_cached = None


def load():
    """Import reference modules; returns a SimpleNamespace of the hot-path callables."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("reference not mounted at %s" % REF_ROOT)
    sys.dont_write_bytecode = True  # the mount is read-only
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if not hasattr(np, "int"):
        np.int = int  # noqa: NPY001 - shim for the reference's NumPy-1.18-era code

    from oracle import librosa_np

    def _nope(*a, **k):
        raise RuntimeError("stubbed third-party function called")

    _stub("soundfile", read=_nope, write=_nope)
    _stub("ipdb", set_trace=_nope)
    mpl = _stub("matplotlib", use=lambda *a, **k: None)
    mpl.pyplot = _stub("matplotlib.pyplot")
    mpl.patches = _stub("matplotlib.patches", Rectangle=object, Circle=object)
    ac = _stub("acoustics")
    from oracle import octave_np
    ac.signal = _stub("acoustics.signal", OctaveBand=octave_np.OctaveBand)   # restated third party
    me = _stub("mir_eval")
    me.separation = _stub("mir_eval.separation", bss_eval_sources=_nope)
    ps = _stub("pystoi")
    ps.stoi = _stub("pystoi.stoi", stoi=_nope)
    core = _stub("librosa.core", stft=librosa_np.stft, istft=librosa_np.istft)
    _stub("librosa", core=core, stft=librosa_np.stft, istft=librosa_np.istft,
          pcen=_nope)

    importlib.import_module("disco_theque")
    _stub("disco_theque.dnn.models.heymann", build_heymann=None)
    # break the dnn.utils <-> dnn.models.crnn cycle
    crnn_stub = _stub("disco_theque.dnn.models.crnn", build_crnn=None)
    dnn_utils = importlib.import_module("disco_theque.dnn.utils")
    del sys.modules["disco_theque.dnn.models.crnn"]
    del crnn_stub
    crnn = importlib.import_module("disco_theque.dnn.models.crnn")
    dnn_utils.build_crnn = crnn.build_crnn
    if not hasattr(dnn_utils, "move_to_device"):
        dnn_utils.move_to_device = lambda x, device=None: x

    tango = importlib.import_module("disco_theque.speech_enhancement.tango")
    formulas = importlib.import_module("disco_theque.se_utils.internal_formulas")
    sigproc = importlib.import_module("disco_theque.sigproc_utils")
    se_utils = importlib.import_module("disco_theque.speech_enhancement.utils")
    metrics = importlib.import_module("disco_theque.metrics")

    ns = types.SimpleNamespace(
        tango=tango,
        offline_tango=tango.offline_tango,
        concatenate_signals=tango.concatenate_signals,
        get_z_for_mask=tango.get_z_for_mask,
        get_mask=tango.get_mask,
        reshape_mask=tango.reshape_mask,
        intern_filter=formulas.intern_filter,
        spatial_correlation_matrix=formulas.spatial_correlation_matrix,
        tf_mask=dnn_utils.tf_mask,
        tf_mask_sigproc=sigproc.tf_mask,
        vad_oracle_batch=sigproc.vad_oracle_batch,
        prepare_data=se_utils.prepare_data,
        build_crnn=crnn.build_crnn,
        metrics=metrics,
        third_octave_filterbank=sigproc.third_octave_filterbank,
    )
    _cached = ns
    return ns


def run_offline_tango(y, s, n, vads=("irm1", "irm1"), mods=(None, None),
                      mask_for_z="local", z_sigs="zs_hat"):
    """Call the reference's offline_tango (tango.py:252) for K = len(y) nodes, ref mic 0."""
    ref = load()
    ref.tango.ref_mics = [0] * len(y)   # module global read at call time (tango.py:338)
    return ref.offline_tango(y, s, n, list(vads), list(mods), mask_for_z, z_sigs)
