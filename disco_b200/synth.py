"""Deterministic synthetic multi-node, multi-mic noisy-speech-like signals.

No corpora are available, so the benchmark and the parity tests use a spatially
coherent target (one Gaussian source convolved with a decaying random 32-tap FIR
per microphone) plus spatially white noise.  White-only inputs would understate
the conditioning problems of the per-bin GEVD.
"""
import numpy as np


def make_utterance(seed, n_nodes, n_ch, length, taps=32, src_std=0.1, noise_std=0.05, gate_period=0):
    """Returns (y, s, n): float32 arrays of shape (n_nodes, n_ch, length), y = s + n.

    gate_period > 0 switches the source off for the last 40 % of every period (speech
    pauses, so that an energy VAD has something to detect)."""
    rng = np.random.default_rng(1234 + int(seed))
    src = (src_std * rng.standard_normal(length + 2 * taps)).astype(np.float64)
    if gate_period:
        src *= (np.arange(len(src)) % gate_period) < 0.6 * gate_period
    decay = np.exp(-np.arange(taps) / 6.0)
    s = np.empty((n_nodes, n_ch, length), np.float32)
    n = np.empty((n_nodes, n_ch, length), np.float32)
    for k in range(n_nodes):
        for c in range(n_ch):
            h = rng.standard_normal(taps) * decay
            s[k, c] = np.convolve(src, h)[taps:taps + length].astype(np.float32)
            n[k, c] = (noise_std * rng.standard_normal(length)).astype(np.float32)
    y = (s + n).astype(np.float32)
    return y, s, n


def make_batch(n_utt, n_nodes, n_ch, length, seed0=0):
    """Batch of utterances: (y, s, n) float32 of shape (n_utt, n_nodes, n_ch, length)."""
    ys, ss, ns = [], [], []
    for b in range(n_utt):
        y, s, n = make_utterance(seed0 + b, n_nodes, n_ch, length)
        ys.append(y), ss.append(s), ns.append(n)
    return np.stack(ys), np.stack(ss), np.stack(ns)
