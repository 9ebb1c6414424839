"""End-to-end parity of the two-step Tango MWF against the reference's own outputs (tests/golden,
produced by the unmodified reference through oracle/ref_shim.py) and the float64 oracle."""
import numpy as np
import pytest
import torch

from conftest import load_golden, record_parity, rel_l2_mag
from oracle.make_golden import NAMES, TANGO_CASES, case_inputs, digest

pytestmark = pytest.mark.gpu

# north-star tolerance: <= 1e-5 relative on the beamformed STFT magnitudes, per (utterance, node).
# the reference's own single-precision LAPACK path can sit further than that from the
# exact mathematics on ill-conditioned inputs (e.g. 9.9e-5 on tango_k2c4_irm2_iam1); a result then also
# passes when it is at least as close to the float64 oracle as the reference itself is (+1e-6).
TOL = 1e-5     # == conftest.TOL


def f64_truth(name, y, s, n, vads, mfz):
    """float64 yardstick: oracle/tango_f64.py (Cholesky-whitened eigh) where it restates the mode, otherwise the
    reference's own algorithm evaluated in double precision (oracle/tango_np.py double=True: complex128 spectra and
    SCMs, LAPACK zggev).  None for VAD / binary masks (their float32 mask inputs are part of the fixture)."""
    from oracle import librosa_np, tango_f64, tango_np
    if any(v[:3] not in ("irm", "iam") for v in vads):
        return None
    chans = [len(c) for c in y]
    if len(set(chans)) != 1 or mfz not in ("local", "distant"):
        res = tango_np.offline_tango(y, s, n, vads=vads, mask_for_z=mfz, granularity="bin", double=True)
        return dict(zip(NAMES, res))
    K = len(y)
    S = [librosa_np.stft(np.asarray(s[k][0])) for k in range(K)]
    N = [librosa_np.stft(np.asarray(n[k][0])) for k in range(K)]
    mz = np.array([tango_np.tf_mask(S[k], N[k], vads[0]) for k in range(K)])
    mw = np.array([tango_np.tf_mask(S[k], N[k], vads[1]) for k in range(K)])
    return tango_f64.offline_tango(np.array(y), np.array(s), np.array(n), masks=(mz, mw), mask_for_z=mfz)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


GPU_CASES = sorted(TANGO_CASES)
# Binary masks make R_nn exactly singular in bins where the mask is 1 in every frame; the reference
# then depends on LAPACK returning inf/NaN generalised eigenvalues (internal_formulas.py:59-62), which
# no other arithmetic can reproduce.  For those cases only the masks are compared.
MASKS_ONLY = {"tango_k2c2_ibm1"}


@pytest.mark.parametrize("name", GPU_CASES)
def test_offline_tango_matches_reference(dev, name):
    from disco_b200.tango import offline_tango
    seed, chans, length, vads, mfz, keep = TANGO_CASES[name]
    g = load_golden(name)
    y, s, n = case_inputs(seed, chans, length, vads)
    assert digest(y) == str(g["input_sha256"])
    res = offline_tango(y, s, n, list(vads), [None, None], mfz)
    assert len(res) == 9
    truth = None
    for nm, val in zip(NAMES, res):
        assert len(val) == len(chans)
        for k in range(len(chans)):
            key = "%s_%d" % (nm, k)
            if key not in g:
                continue
            ref, got = g[key], np.asarray(val[k])
            assert got.shape == ref.shape, key
            if not nm.startswith("mask") and name in MASKS_ONLY:
                assert np.all(np.isfinite(got)), key        # diagonal loading keeps our output finite
                continue
            if nm.startswith("mask"):
                if ref.dtype == bool:
                    assert got.dtype == bool and np.mean(got != ref) < 1e-3, key
                else:
                    # the masks inherit the float32-FFT rounding of |S|, |N| (the reference's STFT is
                    # computed in float64 and rounded): a few 1e-7 relative on xi, more where |N| is tiny
                    # ('iam' = |s| / |s + n| is unbounded and ill-conditioned where s + n cancels: judged on
                    # the 99.5th percentile of the relative deviation instead of the maximum)
                    dev_ = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
                    assert np.percentile(dev_, 99.5) < 5e-6, key
                    if "iam" not in vads[NAMES.index(nm) - 7]:
                        assert np.max(dev_) < 5e-6, key
            else:
                assert got.dtype == np.complex64
                err = rel_l2_mag(got, ref)
                if truth is None:
                    truth = f64_truth(name, y, s, n, vads, mfz) or {}
                ours = theirs = None
                if nm in truth:
                    ours, theirs = rel_l2_mag(got, truth[nm][k]), rel_l2_mag(ref, truth[nm][k])
                ok = record_parity(name, nm, k, err_ref=err, err_f64=ours, ref_f64=theirs,
                                   note="reference output (tests/golden)")
                assert ok, (key, err, ours, theirs)


def test_tango_batched_matches_f64_and_is_batch_invariant(dev):
    """B=3 utterances x K=4 nodes x C=4 mics: each (utterance, node) within TOL of the float64 oracle,
    and identical to processing the utterance alone (no cross-talk between batch entries)."""
    from disco_b200.synth import make_batch
    from disco_b200.tango import tango_batched
    from oracle import tango_f64
    B, K, C, L = 3, 4, 4, 12000
    y, s, n = make_batch(B, K, C, L, seed0=50)
    yd, sd, nd = (torch.from_numpy(a).to(dev) for a in (y, s, n))
    out = tango_batched(yd, sd, nd)
    solo = tango_batched(yd[1:2], sd[1:2], nd[1:2])
    assert torch.equal(out["yf"][1], solo["yf"][0])
    for b in range(B):
        ref = tango_f64.offline_tango(y[b], s[b], n[b])
        for k in range(K):
            for nm in ("yf", "z_y", "sf", "nf", "z_s", "z_n"):
                got = out[nm][b, k].cpu().numpy()
                assert rel_l2_mag(got, ref[nm][k]) < TOL, (nm, b, k, rel_l2_mag(got, ref[nm][k]))


def test_tango_external_masks_deployment_mode(dev):
    """DNN-style deployment: only y and device-resident masks (frame-major), no clean components."""
    from disco_b200.synth import make_batch
    from disco_b200.tango import tango_batched
    from oracle import tango_f64
    B, K, C, L = 2, 2, 4, 16000
    y, _, _ = make_batch(B, K, C, L, seed0=70)
    T, F = 1 + L // 256, 257
    rng = np.random.default_rng(3)
    mz = rng.uniform(0.05, 0.95, size=(B, K, T, F)).astype(np.float32)
    mw = rng.uniform(0.05, 0.95, size=(B, K, T, F)).astype(np.float32)
    out = tango_batched(torch.from_numpy(y).to(dev), masks=(torch.from_numpy(mz).to(dev), torch.from_numpy(mw).to(dev)))
    assert set(out) == {"yf", "z_y", "zn", "masks_z", "mask_w"}
    for b in range(B):
        ref = tango_f64.offline_tango(y[b], masks=(mz[b].transpose(0, 2, 1), mw[b].transpose(0, 2, 1)))
        for k in range(K):
            assert rel_l2_mag(out["yf"][b, k].cpu().numpy(), ref["yf"][k]) < TOL
            assert rel_l2_mag(out["zn"][b, k].cpu().numpy(), ref["zn"][k]) < TOL


def test_reference_error_behaviour(dev):
    from disco_b200.tango import offline_tango
    y, s, n = case_inputs(0, [2, 2], 4000)
    with pytest.raises(TypeError):
        offline_tango(y, s, n, ["irm1", "irm1"], [None, None], None)     # reference tango.py:343
    with pytest.raises(ValueError):
        offline_tango(y, s, n, ["foo1", "irm1"], [None, None], "local")   # reference tango.py:223


def test_cuda_graph_plan_matches_eager(dev):
    """TangoGraph (captured pipeline, 2 batch chunks on parallel branches) == eager tango_batched, and
    replays with new inputs give the new results."""
    from disco_b200.plan import TangoGraph
    from disco_b200.synth import make_batch
    from disco_b200.tango import tango_batched
    B, K, C, L = 5, 1, 4, 20000
    T, F = 1 + L // 256, 257
    plan = TangoGraph(B, K, C, L, chunks=2, device=dev)
    for seed in (90, 91):
        y, _, _ = make_batch(B, K, C, L, seed0=seed)
        g = torch.Generator().manual_seed(seed)
        mz, mw = torch.rand((B, K, T, F), generator=g), torch.rand((B, K, T, F), generator=g)
        plan.load(torch.from_numpy(y).pin_memory(), mz.pin_memory(), mw.pin_memory())
        plan.run()
        got = plan.output("yf")
        ref = tango_batched(torch.from_numpy(y).to(dev), masks=(mz.to(dev), mw.to(dev)), out_layout="TF",
                            diagnostics=False)
        assert torch.equal(got, ref["yf"])
        host = torch.empty((B, K, T, F), dtype=torch.complex64).pin_memory()
        plan.store("yf", host)
        torch.cuda.synchronize()
        assert torch.equal(host, ref["yf"].cpu())


@pytest.mark.parametrize("n_fft", [256, 512, 1024])
@pytest.mark.parametrize("K,C", [(1, 8), (2, 2), (1, 3)])
def test_tango_shapes_of_the_stft_sweep_config(dev, n_fft, K, C):
    """BASELINE configs[3]: STFT-length sweep 256/512/1024 with 8 mics (and small-array variants):
    each (utterance, node) within TOL of the float64 oracle, masks supplied on device."""
    from disco_b200.synth import make_batch
    from disco_b200.tango import tango_batched
    from oracle import tango_f64
    B, L = 2, 24000
    y, _, _ = make_batch(B, K, C, L, seed0=300 + C)
    T, F = 1 + L // (n_fft // 2), n_fft // 2 + 1
    rng = np.random.default_rng(n_fft + C)
    mz = rng.uniform(0.05, 0.95, size=(B, K, T, F)).astype(np.float32)
    mw = rng.uniform(0.05, 0.95, size=(B, K, T, F)).astype(np.float32)
    out = tango_batched(torch.from_numpy(y).to(dev), masks=(torch.from_numpy(mz).to(dev), torch.from_numpy(mw).to(dev)),
                        n_fft=n_fft)
    for b in range(B):
        ref = tango_f64.offline_tango(y[b], masks=(mz[b].transpose(0, 2, 1), mw[b].transpose(0, 2, 1)),
                                      n_fft=n_fft, n_hop=n_fft // 2)
        # against EXACT arithmetic the error grows with the channel count (the conditioning of the 8-channel GEVD
        # amplifies the float32 rounding of the spectra).  The bar stays the reference: where float64 is further
        # than TOL, the reference-precision port (oracle/tango_np: complex64 SCMs, single-precision cggev) must be
        # at least as far from float64 as we are -- recorded per node in the parity table.
        port = None
        for k in range(K):
            for nm in ("yf", "z_y"):
                got = out[nm][b, k].cpu().numpy()
                e64 = rel_l2_mag(got, ref[nm][k])
                eref = pe64 = None
                if e64 >= TOL:
                    if port is None:
                        from oracle import tango_np
                        res = tango_np.offline_tango(y[b], y[b], y[b], masks=(mz[b].transpose(0, 2, 1), mw[b].transpose(0, 2, 1)),
                                                     n_fft=n_fft, n_hop=n_fft // 2, granularity="bin")
                        port = {"yf": res[0], "z_y": res[3]}
                    eref, pe64 = rel_l2_mag(got, port[nm][k]), rel_l2_mag(port[nm][k], ref[nm][k])
                assert record_parity("sweep_nfft%d_k%dc%d_b%d" % (n_fft, K, C, b), nm, k, err_ref=eref, err_f64=e64,
                                     ref_f64=pe64, note="reference = fp32 port (oracle/tango_np)"), (nm, k, e64, pe64)


def test_host_pipeline_matches_eager(dev):
    """TangoPipeline (pinned host in, pinned host out, overlapped slices) == eager tango_batched."""
    from disco_b200.plan import TangoPipeline
    from disco_b200.synth import make_batch
    from disco_b200.tango import tango_batched
    B, K, C, L = 7, 2, 2, 16000
    T, F = 1 + L // 256, 257
    pipe = TangoPipeline(B, K, C, L, chunks=3, device=dev)
    y, _, _ = make_batch(B, K, C, L, seed0=120)
    g = torch.Generator().manual_seed(2)
    yh = torch.from_numpy(y).pin_memory()
    mz, mw = torch.rand((B, K, T, F), generator=g).pin_memory(), torch.rand((B, K, T, F), generator=g).pin_memory()
    out = torch.empty((B, K, T, F), dtype=torch.complex64).pin_memory()
    for _ in range(2):
        pipe.process(yh, mz, mw, out)
    torch.cuda.synchronize()
    ref = tango_batched(yh.to(dev), masks=(mz.to(dev), mw.to(dev)), out_layout="TF", diagnostics=False)
    assert torch.equal(out, ref["yf"].cpu())


def test_crnn_pipeline_matches_manual_chain(dev):
    """CrnnTangoPipeline (int16 PCM host -> device -> CRNN masks -> two-mask fused path -> host) == the same chain
    spelled out with the per-utterance mask estimator and tango_batched."""
    from disco_b200 import dnn_mask, ops
    from disco_b200.plan import CrnnTangoPipeline
    from disco_b200.synth import make_batch
    from disco_b200.tango import tango_batched
    B, C, L = 3, 2, 16000
    T, F = 1 + L // 256, 257
    torch.manual_seed(5)
    models = (dnn_mask.CRNN(1, cnn_filters=(4, 6, 6), rnn_units=(8,)), dnn_mask.CRNN(1, cnn_filters=(4, 6, 6), rnn_units=(8,)))
    pipe = CrnnTangoPipeline(B, C, L, chunks=2, device=dev, models=models, exact=True)
    y, _, _ = make_batch(B, 1, C, L, seed0=900)
    pcm = pipe.to_pcm(torch.from_numpy(y))
    out = torch.empty((B, 1, T, F), dtype=torch.complex64).pin_memory()
    for _ in range(2):
        pipe.process(pcm, out)
    torch.cuda.synchronize()
    rep = pipe.report()
    assert rep["slices"] == 2 and rep["crnn_ms_per_slice"] > 0
    yq = (pcm.to(torch.float32) / 32768.0).to(dev)                      # what soundfile would hand the reference
    Yref = ops.stft(yq[:, 0, 0].contiguous())
    mz = torch.stack([dnn_mask.estimate_mask(pipe.models[0], Yref[b].T, None, device=dev) for b in range(B)])[:, None]
    mw = torch.stack([dnn_mask.estimate_mask(pipe.models[1], Yref[b].T, None, device=dev) for b in range(B)])[:, None]
    ref = tango_batched(yq, masks=(mz.contiguous(), mw.contiguous()), out_layout="TF", diagnostics=False)
    for b in range(B):
        assert rel_l2_mag(out[b, 0].numpy(), ref["yf"][b, 0].cpu().numpy()) < 1e-4     # untrained masks: degenerate GEVD
    assert np.all(np.isfinite(out.numpy().view(np.float32)))


def test_time_domain_outputs(dev):
    """post.to_time: one batched iSTFT for all outputs == librosa-style iSTFT of each (oracle); SI-SDR on the
    device == SI-SDR of the float64 oracle pipeline's output."""
    from disco_b200 import post
    from disco_b200.synth import make_batch
    from disco_b200.tango import tango_batched
    from oracle import librosa_np
    B, K, C, L = 2, 2, 4, 24000
    y, s, n = make_batch(B, K, C, L, seed0=400)
    out = tango_batched(*(torch.from_numpy(a).to(dev) for a in (y, s, n)))
    td = post.to_time(out, L)
    assert set(td) == {"yf", "z_y", "sf", "nf", "z_s", "z_n"} and td["yf"].shape == (B, K, L)
    ref = librosa_np.istft(out["yf"][1, 0].cpu().numpy(), length=L)
    assert np.max(np.abs(td["yf"][1, 0].cpu().numpy() - ref)) < 1e-5
    # SI-SDR of the time-domain output == SI-SDR of the float64 oracle's output (the MWF trades distortion for
    # noise reduction, so no sign of the improvement is guaranteed on a synthetic mixture)
    from oracle import tango_f64
    s_ref = torch.from_numpy(s[:, :, 0]).to(dev)
    got = post.si_sdr(s_ref, td["yf"])
    assert got.shape == (B, K)
    o = tango_f64.offline_tango(y[1], s[1], n[1])
    want = post.si_sdr(s[1, 0, 0], librosa_np.istft(np.asarray(o["yf"][0]).astype(np.complex64), length=L))
    assert abs(got[1, 0].item() - want.item()) < 1e-3
