"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restatement of the one class the reference takes from the third-party package `acoustics`
(requirements.txt:1 pins acoustics==0.2.4.post0; absent from /root/reference and from this image):
``acoustics.signal.OctaveBand(center=f, fraction=3)`` as used by
disco_theque/sigproc_utils.py:112 (`third_octave_filterbank`).

Published algorithm (IEC 61260-1:2014, the standard that package implements):
    G      = 10 ** (3 / 10)                      octave frequency ratio (eq. 1), NOT 2
    index  = round(b * ln(f / f_ref) / ln(G))    for an odd bandwidth designator b (here 3), f_ref = 1000 Hz
    centre = f_ref * G ** (index / b)            exact mid-band frequency (eq. 3)
    lower  = centre * G ** (-1 / (2 b)),  upper = centre * G ** (+1 / (2 b))        (eqs. 4, 5)
No reference test pins these numbers: PARITY UNPINNED for the band edges; everything downstream of them
(Butterworth design, filtering, band powers, weighting) is pinned by running the reference's own
metrics.fw_snr / fw_sd with this class injected (oracle/ref_shim.py, tests/golden/metrics_kat.npz).
"""
import numpy as np

G_OCTAVE = 10.0 ** (3.0 / 10.0)
F_REF = 1000.0


class OctaveBand:
    def __init__(self, center=None, fraction=1, reference=F_REF):
        if fraction % 2 != 1:
            raise NotImplementedError("only odd bandwidth designators are restated (the reference uses 3)")
        c = np.atleast_1d(np.asarray(center, dtype=np.float64))
        index = np.round(fraction * np.log(c / reference) / np.log(G_OCTAVE)).astype(np.int16)
        self.fraction = fraction
        self.center = reference * G_OCTAVE ** (index / fraction)
        self.lower = self.center * G_OCTAVE ** (-1.0 / (2.0 * fraction))
        self.upper = self.center * G_OCTAVE ** (+1.0 / (2.0 * fraction))
        self.bandwidth = self.upper - self.lower
