"""Generate tests/golden/*.npz by running the UNMODIFIED reference — TEST INFRASTRUCTURE.

Run in the authoring container (where /root/reference is mounted):

    python -m oracle.make_golden

The reference's own offline_tango (tango.py:252), intern_filter
(internal_formulas.py:31), tf_mask (dnn/utils.py:44), vad_oracle_batch
(sigproc_utils.py:12), spatial_correlation_matrix (internal_formulas.py:84),
concatenate_signals / get_z_for_mask / reshape_mask (tango.py:142-240) are imported
through oracle/ref_shim.py and executed on seeded synthetic inputs
(disco_b200/synth.py).  Only librosa's STFT is restated (oracle/librosa_np.py).
Inputs are regenerated from their seeds by the tests; a checksum guards the generator.
"""
import hashlib
import os

import numpy as np

from disco_b200.synth import make_utterance
from oracle import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TANGO_CASES = {
    # name: (seed, channels per node, L, vads, mask_for_z, outputs kept)
    "tango_k1c2_cfg1": (0, [2], 64000, ("irm1", "irm1"), "local", ("yf", "z_y")),
    "tango_k2c3_local": (1, [3, 3], 8192, ("irm1", "irm1"), "local", None),
    "tango_k1c8_local": (11, [8], 24000, ("irm1", "irm1"), "local", ("yf", "z_y")),
    "tango_k3_ragged_local": (2, [2, 3, 2], 6000, ("irm1", "irm1"), "local", ("yf", "z_y", "zn", "sf")),
    "tango_k3c2_distant": (3, [2, 2, 2], 6000, ("irm1", "irm1"), "distant", ("yf", "z_y", "nf")),
    "tango_k2c4_irm2_iam1": (4, [4, 4], 24000, ("irm2", "iam1"), "local", ("yf", "z_y", "masks_z", "mask_w")),
    "tango_k2c2_ibm1": (5, [2, 2], 5000, ("ibm1", "ibm1"), "local", ("yf", "z_y", "masks_z")),
    "tango_k2c2_ivad": (6, [2, 2], 8192, ("ivad", "ivad"), "local", ("yf", "z_y", "masks_z")),
    # mask_for_z=None raises TypeError in the reference (tango.py:343 `in None`): no fixture.
    "tango_k2c2_previous": (7, [2, 2], 5000, ("irm1", "irm1"), "previous", ("yf", "z_y")),
    "tango_k2c2_compressed": (8, [2, 2], 5000, ("irm1", "irm1"), "compressed", ("yf", "z_y")),
    "tango_k2c2_oracle_refs": (9, [2, 2], 5000, ("irm1", "irm1"), "use_oracle_refs", ("yf", "z_y")),
    "tango_k2c2_oracle_zs": (10, [2, 2], 5000, ("irm1", "irm1"), "use_oracle_zs", ("yf", "z_y")),
    # geometries of BASELINE configs[2] (4 nodes x 4 mics) and configs[4] (8 nodes x 2 mics), short signals
    "tango_k4c4_local": (12, [4, 4, 4, 4], 8000, ("irm1", "irm1"), "local", ("yf", "z_y")),
    "tango_k8c2_local": (13, [2] * 8, 6000, ("irm1", "irm1"), "local", ("yf", "z_y")),
    # single node, two DIFFERENT masks: the two-mask fused route (stft_scm2 + filter_dual) against the reference
    "tango_k1c4_irm1_irm2": (14, [4], 16000, ("irm1", "irm2"), "local", ("yf", "z_y", "zn")),
    "tango_k1c3_iam1_irm1": (15, [3], 7000, ("iam1", "irm1"), "local", ("yf", "z_y", "zn")),
    # ragged channel counts under the other mask_for_z modes (tango.py:396-429)
    "tango_k3_ragged_distant": (16, [2, 3, 2], 6000, ("irm1", "irm1"), "distant", ("yf", "z_y")),
    "tango_k3_ragged_compressed": (17, [3, 1, 2], 6000, ("irm1", "irm2"), "compressed", ("yf", "z_y")),
    # (ragged + 'use_oracle_refs' raises inside the reference: np.array(s_stft) of ragged lists, tango.py:405)
}
NAMES = ("yf", "sf", "nf", "z_y", "z_s", "z_n", "zn", "masks_z", "mask_w")


def case_inputs(seed, chans, length, vads=("irm1", "irm1")):
    """[node][channel] lists as the reference expects (ragged channel counts allowed)."""
    K, cmax = len(chans), max(chans)
    y, s, n = make_utterance(seed, K, cmax, length, gate_period=2000 if "ivad" in vads else 0)
    pick = lambda a: [[a[k, c] for c in range(chans[k])] for k in range(K)]
    return pick(y), pick(s), pick(n)


def digest(y):
    h = hashlib.sha256()
    for node in y:
        for ch in node:
            h.update(np.ascontiguousarray(ch).tobytes())
    return h.hexdigest()


def rand_hpd(rng, d, rank=None, dtype=np.complex64):
    r = d + 2 if rank is None else rank
    a = rng.standard_normal((d, r)) + 1j * rng.standard_normal((d, r))
    return (a @ a.conj().T / r).astype(dtype)


def main(only=None):
    """only: iterable of Tango case names to (re)generate; None regenerates every fixture."""
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shim.load()

    for name, (seed, chans, length, vads, mfz, keep) in TANGO_CASES.items():
        if only is not None and name not in only:
            continue
        y, s, n = case_inputs(seed, chans, length, vads)
        res = ref_shim.run_offline_tango(y, s, n, vads=vads, mask_for_z=mfz)
        blob = {"input_sha256": np.array(digest(y))}
        for nm, val in zip(NAMES, res):
            if keep is not None and nm not in keep:
                continue
            for k, arr in enumerate(val):
                arr = np.asarray(arr)
                if np.iscomplexobj(arr):
                    arr = arr.astype(np.complex64)
                elif arr.dtype == np.float64 and nm.startswith("mask"):
                    arr = arr.astype(np.float32) if np.array_equal(arr.astype(np.float32), arr) else arr
                blob["%s_%d" % (nm, k)] = arr
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)
        print(name, {k: v.shape for k, v in blob.items() if k.endswith("_0")})

    if only is not None:
        return
    # ---- intern_filter known answers (complex64 and complex128 inputs, all three types)
    rng = np.random.default_rng(42)
    blob = {}
    i = 0
    for d in (2, 3, 4, 7, 9):
        for dt in (np.complex64, np.complex128):
            Rxx = rand_hpd(rng, d, rank=None, dtype=dt) + 4 * rand_hpd(rng, d, rank=1, dtype=dt)
            Rnn = rand_hpd(rng, d, dtype=dt)
            for typ, rank, mu in (("gevd", 1, 1), ("gevd", 2, 1), ("gevd", 1, 3.5), ("gevd", "full", 1),
                                  ("mwf", None, 1), ("r1-mwf", None, 1), ("r1-mwf", None, 2.0)):
                if rank is None:
                    W, (t1, _) = ref.intern_filter(Rxx, Rnn, mu=mu, type=typ)
                else:
                    if typ == "gevd" and rank != "full" and rank > d:
                        continue
                    W, (t1, _) = ref.intern_filter(Rxx, Rnn, mu=mu, type=typ, rank=rank)
                blob["Rxx_%d" % i], blob["Rnn_%d" % i] = Rxx, Rnn
                blob["W_%d" % i], blob["t1_%d" % i] = np.asarray(W), np.asarray(t1)
                blob["cfg_%d" % i] = np.array("%s|%s|%s" % (typ, rank, mu))
                i += 1
    blob["count"] = np.array(i)
    np.savez_compressed(os.path.join(OUT, "intern_filter_kat.npz"), **blob)
    print("intern_filter KATs:", i)

    # ---- masks, VAD, recursive SCM, list helpers
    rng = np.random.default_rng(7)
    sa = (rng.standard_normal((17, 23)) + 1j * rng.standard_normal((17, 23))).astype(np.complex64)
    na = (rng.standard_normal((17, 23)) + 1j * rng.standard_normal((17, 23))).astype(np.complex64)
    na[3, 4] = 0
    sa[5, 6] = 0
    blob = {"s": sa, "n": na}
    for typ in ("irm1", "irm2", "ibm1", "ibm2", "iam1", "iam2"):
        blob["dnn_" + typ] = ref.tf_mask(sa, na, type=typ)
        blob["sig_" + typ] = ref.tf_mask_sigproc(sa, na, type=typ)
    blob["dnn_ibm1_thr3"] = ref.tf_mask(sa, na, type="ibm1", bin_thr=3)
    x = (0.1 * rng.standard_normal(5000) * (np.arange(5000) % 2000 < 900)).astype(np.float32)
    blob["vad_x"] = x
    blob["vad_default"] = ref.vad_oracle_batch(x)
    blob["vad_256_128"] = ref.vad_oracle_batch(x, win_len=256, win_hop=128, thr=0.01, rat=4)
    R0 = rand_hpd(rng, 4, dtype=np.complex128)
    xv = rng.standard_normal(4) + 1j * rng.standard_normal(4)
    blob["scm_R0"], blob["scm_x"] = R0, xv
    blob["scm_plain"] = ref.spatial_correlation_matrix(R0, xv)
    blob["scm_masked"] = ref.spatial_correlation_matrix(R0, xv, lambda_cor=0.9, M=0.3)
    ys = [(rng.standard_normal((2, 5, 6)) + 0j).astype(np.complex64) for _ in range(3)]
    zs = [(rng.standard_normal((5, 6)) * 1j).astype(np.complex64) for _ in range(3)]
    zn_ = [(rng.standard_normal((5, 6)) + 0j).astype(np.complex64) for _ in range(3)]
    mk = rng.uniform(size=(5, 6)).astype(np.float32)
    blob["cat_y"], blob["cat_z"], blob["cat_zn"], blob["cat_m"] = np.array(ys), np.array(zs), np.array(zn_), mk
    for k in range(3):
        blob["cat_%d" % k] = ref.concatenate_signals(ys, zs, k)
        blob["catm_%d" % k] = ref.concatenate_signals(ys, zs, k, mk)
        blob["zmask_zs_%d" % k] = ref.get_z_for_mask(zs, zn_, k, 3, "zs_hat")
        blob["zmask_zn_%d" % k] = ref.get_z_for_mask(zs, zn_, k, 3, "zn_hat")
        blob["zmask_both_%d" % k] = ref.get_z_for_mask(zs, zn_, k, 3, ["zs_hat", "zn_hat"])
    mstack = rng.uniform(size=(9, 15, 11)).astype(np.float32)
    blob["reshape_in"] = mstack
    blob["reshape_last"] = ref.reshape_mask(mstack, "last")
    blob["reshape_mid"] = ref.reshape_mask(mstack, "mid")
    np.savez_compressed(os.path.join(OUT, "helpers_kat.npz"), **blob)
    print("helpers KATs written")

    # ---- DNN-mask feeder: the reference's prepare_data + CRNN forward + reshape_mask on a tiny CRNN
    import torch
    torch.manual_seed(3)
    rng = np.random.default_rng(5)
    blob = {}
    for n_ch in (1, 3):
        model, _ = ref.build_crnn((n_ch, 21, 257), (4, 6, 6), (3, 3, 3), (1, 1, 1), [(1, 4)] * 3, (None,) * 3, [8],
                                  "GRU", 257, conv_padding=[(0, 1)] * 3)
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.data.uniform_(0.5, 1.5)
                m.bias.data.normal_(0, 0.1)
        model.eval()
        T = 37
        ysp = (rng.standard_normal((257, T)) + 1j * rng.standard_normal((257, T))).astype(np.complex64)
        zsp = [(rng.standard_normal((257, T)) + 1j * rng.standard_normal((257, T))).astype(np.complex64)
               for _ in range(n_ch - 1)]
        lost = int(21 - model.get_loss_frames("last")[-1][-1])
        tag = "c%d_" % n_ch
        for k, v in model.state_dict().items():
            blob[tag + "sd_" + k] = v.numpy()
        blob[tag + "y"] = ysp
        if zsp:
            blob[tag + "z"] = np.array(zsp)
        for ftp in ("mid", "last"):
            xin = ref.prepare_data(ysp, True, z_data=zsp or None, win_len=21, win_hop=1, frame_to_pred=ftp,
                                   frames_lost=lost)
            with torch.no_grad():
                blob[tag + "mask_" + ftp] = ref.reshape_mask(model(xin).numpy(), ftp)
        blob[tag + "windows_mid_head"] = ref.prepare_data(ysp, True, z_data=zsp or None, win_len=21, win_hop=1,
                                                          frame_to_pred="mid", frames_lost=lost)[:3].numpy()
    np.savez_compressed(os.path.join(OUT, "crnn_kat.npz"), **blob)
    print("CRNN KATs written")
    metrics_kat(ref)
    post_generator_kat(ref)
    online_kat(ref)


def online_inputs(seed=31, D=3, F=129, T=37):
    """Seeded spectra / mask of the online known-answer case (shared with the tests)."""
    rng = np.random.default_rng(seed)
    src = rng.standard_normal((F, T)) + 1j * rng.standard_normal((F, T))
    steer = rng.standard_normal((D, F, 1)) + 1j * rng.standard_normal((D, F, 1))
    X = (steer * src[None] + 0.3 * (rng.standard_normal((D, F, T)) + 1j * rng.standard_normal((D, F, T)))).astype(np.complex64)
    mask = rng.uniform(0.05, 0.95, size=(F, T)).astype(np.float32)
    return X, mask


def online_kat(ref):
    """Recursive MWF step as the per-frame composition of the REFERENCE's spatial_correlation_matrix and
    intern_filter (oracle/online_np.py drives them): z, filters and smoothed SCM snapshots."""
    from oracle import online_np
    X, mask = online_inputs()
    blob = {}
    for tag, kw in (("p8l1", dict(block=8, lag=1)), ("p5l0", dict(block=5, lag=0)),
                    ("p8l1pow1", dict(block=8, lag=1, power=1, lambda_cor=0.9))):
        z, W, Rs, Rn = online_np.online_mwf(X, mask, ref.spatial_correlation_matrix, ref.intern_filter, **kw)
        blob[tag + "_z"], blob[tag + "_W"] = z.astype(np.complex64), W.astype(np.complex64)
        blob[tag + "_Rss"], blob[tag + "_Rnn"] = Rs[-2:].astype(np.complex64), Rn[-2:].astype(np.complex64)
    np.savez_compressed(os.path.join(OUT, "online_kat.npz"), **blob)
    print("online KATs written")


def make_post_dataset(root, seed=21, rirs=(1, 2), lengths=(6000, 5200), noise="fs", scene="living", case="train"):
    """A tiny on-disk data set in the layout PostGenerator reads (post_generator.py:86-97): 16 convolved target
    and 16 convolved noise channels per RIR, 16-bit PCM.  Shared by the golden generator and the tests."""
    from disco_b200 import wav_io          # the WAV codec restatement (soundfile is absent)
    rng = np.random.default_rng(seed)
    base = os.path.join(root, scene, case, "wav_original", "cnv")
    os.makedirs(os.path.join(base, "target"), exist_ok=True)
    os.makedirs(os.path.join(base, "noise"), exist_ok=True)
    for rir, L in zip(rirs, lengths):
        src = rng.standard_normal(L + 31) * 0.1
        for ch in range(1, 17):
            h = rng.standard_normal(32) * np.exp(-np.arange(32) / 6.0)
            wav_io.write(os.path.join(base, "target", "%d_S-1_Ch-%d.wav" % (rir, ch)), np.convolve(src, h, "valid") * 0.5, 16000)
            Ln = L - 400 if rir % 2 == 0 else L            # a shorter noise file exercises the zero padding
            wav_io.write(os.path.join(base, "noise", "%d_S-2_%s_Ch-%d.wav" % (rir, noise, ch)),
                         rng.standard_normal(Ln) * 0.05, 16000)


def post_generator_kat(ref):
    """Outputs of the reference's PostGenerator.post_process (dataset_utils/post_generator.py) on the tiny data set,
    with `soundfile` backed by disco_b200.wav_io (the only substitution)."""
    import importlib
    import sys
    import tempfile
    from disco_b200 import wav_io
    sfm = sys.modules["soundfile"]
    sfm.read, sfm.write = wav_io.read, wav_io.write
    pg = importlib.import_module("disco_theque.dataset_utils.post_generator")
    blob = {}
    with tempfile.TemporaryDirectory() as root:
        make_post_dataset(root)
        np.random.seed(7)
        gen = pg.PostGenerator(1, 2, "living", "fs", [0, 6], root, n_samples=[10, 2, 2])
        gen.post_process()
        out = os.path.join(root, "living", "train")
        for rir in (1, 2):
            blob["snr_%d" % rir] = np.load(os.path.join(out, "log", "snrs", "dry", "0-6", "%d_fs.npy" % rir))
            digest_m, digest_s = [], []
            for ch in range(1, 17):
                m = np.load(os.path.join(out, "mask_processed", "0-6", "%d_fs_Ch-%d.npy" % (rir, ch)))
                x = np.load(os.path.join(out, "stft_processed", "raw", "0-6", "mixture", "%d_fs_Ch-%d.npy" % (rir, ch)))
                digest_m.append(float(np.sum(m.astype(np.float64))))
                digest_s.append(float(np.sum(np.abs(x).astype(np.float64))))
                if ch in (1, 7, 16):
                    blob["mask_%d_%d" % (rir, ch)] = m
                if ch == 7:
                    blob["mix_%d_%d" % (rir, ch)] = x
                    blob["noi_%d_%d" % (rir, ch)] = np.load(os.path.join(out, "stft_processed", "raw", "0-6", "noise", "%d_fs_Ch-%d.npy" % (rir, ch)))
                    blob["tar_%d_%d" % (rir, ch)] = np.load(os.path.join(out, "stft_processed", "raw", "0-6", "target", "%d_Ch-%d.npy" % (rir, ch)))
                    blob["abs_%d_%d" % (rir, ch)] = np.load(os.path.join(out, "stft_processed", "normed", "abs", "0-6", "mixture", "%d_fs_Ch-%d.npy" % (rir, ch)))
                    blob["wavmix_%d_%d" % (rir, ch)] = wav_io.read(os.path.join(out, "wav_processed", "0-6", "mixture", "%d_fs_Ch-%d.wav" % (rir, ch)))[0]
                    blob["wavnoi_%d_%d" % (rir, ch)] = wav_io.read(os.path.join(out, "wav_processed", "0-6", "noise", "%d_fs_Ch-%d.wav" % (rir, ch)))[0]
            blob["mask_sums_%d" % rir], blob["mix_abs_sums_%d" % rir] = np.array(digest_m), np.array(digest_s)
        tree = sorted(os.path.relpath(os.path.join(d, f), out) for d, _, fs in os.walk(out) for f in fs
                      if "wav_original" not in d)
        blob["tree"] = np.array(tree)
    np.savez_compressed(os.path.join(OUT, "post_generator_kat.npz"), **blob)
    print("PostGenerator KATs written (%d files in the tree)" % len(tree))


def metrics_kat(ref):
    """Time-domain metrics of the reference (disco_theque/metrics.py) on seeded signals: fw_snr, fw_sd (with
    the restated third-party OctaveBand injected, oracle/octave_np.py), snr, delta_snr, sd, si_sdr, and the
    filter-bank coefficients of sigproc_utils.third_octave_filterbank."""
    rng = np.random.default_rng(11)
    L, fs = 12000, 16000
    colour = np.exp(-np.arange(24) / 4.0) * rng.standard_normal(24)
    s = np.stack([np.convolve(rng.standard_normal(L + 23), colour, "valid") * g for g in (0.1, 0.02, 0.3)])
    n = rng.standard_normal((3, L)) * np.array([[0.05], [0.05], [0.01]])
    s[1, :700] = 0.0                      # leading zeros: exactly-zero filter outputs are excluded (metrics.py:100)
    n[2, :300] = 0.0
    s, n = s.astype(np.float32), n.astype(np.float32)
    est = (0.8 * s + 0.3 * n).astype(np.float32)
    vad = (np.abs(s) > 0.5 * np.std(s, axis=1, keepdims=True)).astype(np.float64)
    m = ref.metrics
    blob = {"s": s, "n": n, "est": est, "vad": vad, "fs": fs}
    for i in range(3):
        fq, mean, F = m.fw_snr(s[i], n[i], fs)
        blob["fw_snr_fq_%d" % i], blob["fw_snr_mean_%d" % i] = fq, mean
        fq, mean, _ = m.fw_snr(s[i], n[i], fs, vad_tar=vad[i], vad_noi=vad[i])
        blob["fw_snr_vad_fq_%d" % i], blob["fw_snr_vad_mean_%d" % i] = fq, mean
        fq, mean, _ = m.fw_snr(s[i], n[i], fs, clipping=0, db=False)
        blob["fw_snr_lin_fq_%d" % i], blob["fw_snr_lin_mean_%d" % i] = fq, mean
        fq, mean, _ = m.fw_sd(est[i], s[i], fs)
        blob["fw_sd_fq_%d" % i], blob["fw_sd_mean_%d" % i] = fq, mean
        blob["snr_%d" % i] = m.snr(s[i], n[i])
        blob["sd_%d" % i] = m.sd(est[i], s[i])
        blob["delta_snr_%d" % i] = m.delta_snr(0.8 * s[i], 0.3 * n[i], s[i], n[i])
        blob["si_sdr_%d" % i] = m.si_sdr(s[i].astype(np.float64), est[i].astype(np.float64))
    blob["F"] = F
    b, a = ref.third_octave_filterbank(F, fs, order=4)
    blob["bank_b4"], blob["bank_a4"] = b, a
    fq, mean, F8 = m.fw_snr(s[0], n[0], 8000)
    blob["fw_snr8k_fq"], blob["fw_snr8k_mean"], blob["F8k"] = fq, mean, F8
    np.savez_compressed(os.path.join(OUT, "metrics_kat.npz"), **blob)
    print("metrics KATs written")


if __name__ == "__main__":
    import sys
    main(only=set(sys.argv[1:]) or None)      # python -m oracle.make_golden [case names ...]
