"""Dataset post-processing on the device: the bulk producer of the training data the
mask estimators consume.  Mirrors

    PostGenerator               disco_theque/dataset_utils/post_generator.py:9-166
    get_z_signals.main          disco_theque/speech_enhancement/get_z_signals.py:318-359 (saving of z_s_hat / z_n_hat)

with the reference's constructor arguments, directory layout and file names
(`{wav,stft}_processed/...`, `mask_processed/...`, `log/snrs/dry/...`, `stft_z/...`), so `DiscoDataset`
(dnn/data/datasets.py:80) reads the result unchanged.  What changes is the arithmetic path: per RIR the
reference runs 48 single-signal librosa STFTs and 16 NumPy mask evaluations; here the 3 x 16 signals are
ONE batched STFT launch and the 16 masks one elementwise kernel, and several RIRs can share a launch
(`batch`).  Files are written by the host exactly as the reference writes them (np.save of (F, T) arrays).
"""
import glob
import os

import numpy as np
import torch

from . import ops, wav_io


class PostGenerator:
    """Reference signature: post_generator.py:17-55."""

    def __init__(self, rir_start, nb_rir, scene, noise, snr_range, path_to_dataset, n_fft=512, n_hop=256,
                 mask_type="irm1", save_target=True, n_samples=None, device="cuda", batch=8):
        if n_hop * 2 != n_fft:
            raise ValueError("the STFT kernels use hop = n_fft / 2 (the reference's N_FFT = 512, N_HOP = 256)")
        self.rir_start, self.nb_rir = rir_start, nb_rir
        self.save_target = save_target
        self.scene, self.noise = scene, noise
        self.snr_range = np.array(snr_range)
        self.snr_out = np.zeros((nb_rir, 1))
        self.path_dataset = path_to_dataset
        self.n_fft, self.n_hop, self.mask_type = n_fft, n_hop, mask_type
        self.snr_dir = self.get_directory_name()
        n_train, n_val, n_test = n_samples if n_samples is not None else (10000, 1000, 1000)
        self.n_samples = np.cumsum([n_train, n_val, n_test])
        self.case = self.get_dset()
        self.fs = 16000
        self.ch_per_node = [4, 4, 4, 4]
        self.n_ch = sum(self.ch_per_node)
        self.n_nodes = len(self.ch_per_node)
        self.device, self.batch = torch.device(device), max(1, int(batch))

    def get_dset(self):
        """post_generator.py:57-63."""
        assert 0 < self.rir_start < self.n_samples[-1], \
            "rir should be between 1 and {}".format(str(self.n_samples[-1]))
        first = np.where(self.rir_start < self.n_samples)[0][0]
        assert self.rir_start + self.nb_rir < self.n_samples[first], \
            "First and last RIRs do not belong to the same set."
        return ["train", "val", "test"][first]

    def get_directory_name(self):
        return "{}-{}".format(str(self.snr_range[0]), str(self.snr_range[1]))

    # ------------------------------------------------------------------ I/O (host)
    def get_sig_lists(self, rir):
        """post_generator.py:86-97."""
        root = os.path.join(self.path_dataset, self.scene, self.case, "wav_original", "cnv")
        key = lambda x: int(x.split("_Ch-")[-1].split(".wav")[0])
        tar = sorted(glob.glob(os.path.join(root, "target", "") + str(rir) + "_S-1_Ch-*.wav"), key=key)
        noi = sorted(glob.glob(os.path.join(root, "noise", "") + str(rir) + "_S-2_" + self.noise + "_Ch-*.wav"), key=key)
        return tar, [noi]

    def _done(self, rir):
        return os.path.isfile(os.path.join(self.path_dataset, self.scene, self.case, "log", "snrs", "dry", self.snr_dir, "")
                              + "{}_{}.npy".format(str(rir), self.noise))

    def load_sigs(self, tar_list, noi_list):
        """The loading half of mix_sigs (post_generator.py:99-118): float32 arrays [n_ch, L]; noise zero-padded to
        the target length like `noi_seg[:len(noi)] += noi`; a noise file LONGER than the target raises ValueError, as
        that NumPy statement does in the reference."""
        tars = [wav_io.read(tar_list[ch], dtype="float32")[0] for ch in range(self.n_ch)]
        L = len(tars[0])
        nois = np.zeros((self.n_ch, L), dtype=np.float32)
        for ch in range(self.n_ch):
            noi = wav_io.read(noi_list[0][ch], dtype="float32")[0]
            if len(noi) > L:
                raise ValueError("noise file %s has %d samples, more than the target's %d" % (noi_list[0][ch], len(noi), L))
            nois[ch, :len(noi)] = noi
        return np.stack(tars), nois

    # ------------------------------------------------------------------ compute (device)
    def process_batch(self, tars, nois, snrs):
        """tars, nois: lists of [n_ch, L] float32 arrays (same L within the batch), snrs: one SNR per item.
        Returns per item (tar, noi, mix) float64 time signals and (S, N, M, mask) in the (F, T) layout."""
        dev = self.device
        t = torch.from_numpy(np.stack(tars)).to(dev)                           # [R, n_ch, L]
        g = torch.tensor([10 ** (-snr / 20) for snr in snrs], dtype=torch.float32, device=dev).view(-1, 1, 1)
        n = torch.from_numpy(np.stack(nois)).to(dev) * g                       # float32 product, as NumPy does
        mix64 = t.double() + n.double()                                        # the reference mixes in float64
        sig = torch.stack([t, n, mix64.float()])                               # [3, R, n_ch, L]
        spec = ops.stft(sig.contiguous(), self.n_fft)                          # ONE launch: [3, R, n_ch, T, F]
        mask = ops.tf_mask(spec[0], spec[1], self.mask_type)                   # [R, n_ch, T, F]
        spec_ft = ops.transpose_last2(spec)                                    # (F, T) like the reference's files
        mask_ft = ops.transpose_last2(mask).cpu().numpy()
        if self.mask_type.startswith("ibm"):
            mask_ft = mask_ft.astype(bool)                                     # the reference's binary mask is boolean
        return (t.double().cpu().numpy(), n.double().cpu().numpy(), mix64.cpu().numpy(),
                spec_ft.cpu().numpy(), mask_ft)

    def post_process(self):
        """post_generator.py:70-84, RIRs grouped into batches of equal length."""
        path_out = os.path.join(self.path_dataset, self.scene, self.case)
        os.makedirs(os.path.join(path_out, "log", "snrs", "dry", self.snr_dir), exist_ok=True)
        pending = []
        for rir in range(self.rir_start, self.rir_start + self.nb_rir):
            if self._done(rir):
                print("{} already processed".format(str(rir)))
                continue
            tar_list, noi_list = self.get_sig_lists(rir)
            tars, nois = self.load_sigs(tar_list, noi_list)
            # one random SNR per RIR, drawn in RIR order like the reference (post_generator.py:103)
            snr = self.snr_range[0] + (self.snr_range[1] - self.snr_range[0]) * np.random.random()
            self.snr_out[rir - self.rir_start, :] = snr
            if pending and pending[0][1].shape != tars.shape:      # a batch shares one signal length
                self._flush(pending)
                pending = []
            pending.append((rir, tars, nois, snr))
            if len(pending) == self.batch:
                self._flush(pending)
                pending = []
        self._flush(pending)

    def _flush(self, items):
        if not items:
            return
        s, n, m, spec, masks = self.process_batch([it[1] for it in items], [it[2] for it in items],
                                                  [it[3] for it in items])
        for i, (rir, _, _, _) in enumerate(items):
            self.save_data(s[i], n[i], m[i], spec[0][i], spec[1][i], spec[2][i], masks[i], rir)

    def save_data(self, s, n, m, ss, ns, ms, masks, rir):
        """post_generator.py:136-166: same folders, same file names, same array layouts."""
        path_out = os.path.join(self.path_dataset, self.scene, self.case)
        for folder in [os.path.join("stft_processed", "raw", ""), "wav_processed"]:
            for sub in ["target", "noise", "mixture"]:
                os.makedirs(os.path.join(path_out, folder, self.snr_dir, sub), exist_ok=True)
        os.makedirs(os.path.join(path_out, "stft_processed", "normed", "abs", self.snr_dir, "mixture"), exist_ok=True)
        os.makedirs(os.path.join(path_out, "mask_processed", self.snr_dir), exist_ok=True)
        j = os.path.join
        for i in range(s.shape[0]):
            tag_t = "{}_Ch-{}".format(str(rir), str(i + 1))
            tag_n = "{}_{}_Ch-{}".format(str(rir), self.noise, str(i + 1))
            if self.save_target:
                wav_io.write(j(path_out, "wav_processed", self.snr_dir, "target", "") + tag_t + ".wav", s[i], self.fs)
                np.save(j(path_out, "stft_processed", "raw", self.snr_dir, "target", "") + tag_t, ss[i])
            wav_io.write(j(path_out, "wav_processed", self.snr_dir, "noise", "") + tag_n + ".wav", n[i], self.fs)
            wav_io.write(j(path_out, "wav_processed", self.snr_dir, "mixture", "") + tag_n + ".wav", m[i], self.fs)
            np.save(j(path_out, "stft_processed", "raw", self.snr_dir, "noise", "") + tag_n, ns[i])
            np.save(j(path_out, "stft_processed", "raw", self.snr_dir, "mixture", "") + tag_n, ms[i])
            np.save(j(path_out, "stft_processed", "normed", "abs", self.snr_dir, "mixture", "") + tag_n, abs(ms[i]))
            np.save(j(path_out, "mask_processed", self.snr_dir, "") + tag_n, masks[i])
        np.save(j(path_out, "log", "snrs", "dry", self.snr_dir, "") + "{}_{}".format(str(rir), self.noise),
                self.snr_out[rir - self.rir_start, :])


def save_z_signals(z_sh, z_nh, save_dir_root, dirry, i_rir, noise):
    """The saving half of get_z_signals.main (get_z_signals.py:333-359): per node the compressed signal estimates
    zs_hat / zn_hat, raw complex64 and magnitude, under stft_z/<name>/{raw, normed/abs}/<snr>/{zs_hat, zn_hat}/."""
    for sub in (("raw",), ("normed", "abs")):
        for kind in ("zs_hat", "zn_hat"):
            os.makedirs(os.path.join(save_dir_root, *sub, dirry, kind), exist_ok=True)
    for i_node in range(len(z_sh)):
        tag = "{}_{}_Node-{}".format(str(i_rir), noise, str(i_node + 1))
        for kind, z in (("zs_hat", z_sh[i_node]), ("zn_hat", z_nh[i_node])):
            z = np.asarray(z)
            np.save(os.path.join(save_dir_root, "raw", dirry, kind, "") + tag, z)
            np.save(os.path.join(save_dir_root, "normed", "abs", dirry, kind, "") + tag, abs(z))
