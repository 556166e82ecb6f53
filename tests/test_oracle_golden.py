"""CPU: the oracle (oracle/ref_path.py) against the fixtures produced by the unmodified reference
(tests/golden, oracle/make_golden.py) and against first-principles float64 definitions."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import dft64, ref_path as R


def cfg_of(g):
    sr, n_fft, hop, win, al = (int(v) for v in g["params"])
    return dict(sampling_rate=sr, n_fft=n_fft, hop_length=hop, win_length=win, audio_length=al)


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("name", ["cfg2_small", "default_small"])
def test_stft_istft_match_reference(name):
    g = golden(f"stft_{name}.npz")
    cfg = cfg_of(g)
    for i in range(3):
        wav = torch.from_numpy(g[f"wav{i}"])
        X, mag, ph = R.compute_stft(wav, **cfg)
        y = R.compute_invert_stft(X, **cfg)
        if i == 0:
            assert np.array_equal(X.numpy(), g["X0"])
            assert np.array_equal(mag.numpy(), g["mag0"])
            assert np.array_equal(ph.numpy(), g["phase0"])
        assert np.array_equal(y.numpy(), g[f"istft{i}"])
        assert np.array_equal(R.zero_mean_unit_var_norm(y).numpy(), g[f"norm{i}"])
    X1, _, _ = R.compute_stft(torch.from_numpy(g["wav1d"]), **cfg)
    assert np.array_equal(X1.numpy(), g["X1d"])
    assert np.array_equal(R.compute_invert_stft(X1, **cfg).numpy(), g["istft1d"])


@pytest.mark.parametrize("name", ["cfg2_small", "default_small"])
def test_mask_arithmetic_matches_reference_lines(name):
    g = golden(f"explain_{name}.npz")
    cfg = cfg_of(g)
    wav, mask = torch.from_numpy(g["wav"]), torch.from_numpy(g["mask"])
    _, mag, ph = R.compute_stft(wav, **cfg)
    rel, irr = R.mask_apply_log1p(mag, ph, mask)
    assert np.array_equal(rel.numpy(), g["rel_spec"]) and np.array_equal(irr.numpy(), g["irr_spec"])
    rw, iw = R.explain(wav, mask, mode="log1p", **cfg)
    assert np.array_equal(rw.numpy(), g["rel_wav"]) and np.array_equal(iw.numpy(), g["irr_wav"])
    rn, inn = R.explain(wav, mask, mode="log1p", normalize=True, **cfg)
    assert np.array_equal(rn.numpy(), g["rel_norm"]) and np.array_equal(inn.numpy(), g["irr_norm"])
    lr, li = R.explain(wav, mask, mode="linear", **cfg)
    assert np.array_equal(lr.numpy(), g["lin_rel_wav"]) and np.array_equal(li.numpy(), g["lin_irr_wav"])


def test_metrics_match_reference_functions():
    g = golden("lmac_metrics.npz")
    for tag in ("ka", "rand"):
        p, th, q = (torch.from_numpy(g[f"{tag}_{k}"]) for k in ("p", "theta", "q"))
        assert np.array_equal(R.faithfulness(p, q).numpy(), g[f"{tag}_ff"])
        assert np.array_equal(R.fidelity(th, p).numpy(), g[f"{tag}_fid"])
        assert np.array_equal(R.average_drop(th, p).numpy(), g[f"{tag}_ad"])
        assert np.array_equal(R.average_increase(th, p).numpy(), g[f"{tag}_ai"])
        assert np.array_equal(R.average_gain(th, p).numpy(), g[f"{tag}_ag"])
    # the known answers of SURVEY.md 8(a10)
    np.testing.assert_allclose(g["ka_ff"], [.8, .6, 0, .05], atol=1e-6)
    np.testing.assert_allclose(g["ka_fid"].ravel(), [1, 1, 0, 0])
    np.testing.assert_allclose(g["ka_ad"], [0, 25, 0, 0], atol=1e-4)
    np.testing.assert_allclose(g["ka_ai"], [100, 0, 100, 0])
    np.testing.assert_allclose(g["ka_ag"], [50, 0, 20, 0], atol=1e-4)


def test_logreg_maskhead_tdmask_match_reference():
    g = golden("logreg.npz")
    lg, pr = R.logreg(torch.from_numpy(g["feats"]), torch.from_numpy(g["coef"]), torch.from_numpy(g["intercept"]))
    assert np.array_equal(lg.numpy(), g["logits"]) and np.array_equal(pr.numpy(), g["probs"])
    g = golden("mask_head.npz")
    m = R.mask_head(torch.from_numpy(g["y1"]), torch.from_numpy(g["weight"]), torch.from_numpy(g["bias"]))
    np.testing.assert_allclose(m.numpy(), g["mask"], atol=1e-6)
    g = golden("td_mask.npz")
    m, rel, irr = R.td_mask(torch.from_numpy(g["wave"]).squeeze(0), torch.from_numpy(g["attr"]).squeeze(0))
    assert np.array_equal(m.numpy(), g["mask"]) and np.array_equal(rel.numpy(), g["rel"])
    assert np.array_equal(irr.numpy(), g["irr"])


def test_mel_matches_reference():
    g = golden("mel_default_small.npz")
    sr, n_fft, hop, win, n_mels = (int(v) for v in g["params"])
    mel = R.mel_transform(torch.from_numpy(g["wav"]), sampling_rate=sr, n_fft=n_fft, hop_length=hop,
                          win_length=win, n_mels=n_mels)
    np.testing.assert_allclose(mel.numpy(), g["mel"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("n_fft,hop,win,N", [(64, 20, 64, 300), (128, 40, 80, 500), (64, 16, 64, 257)])
def test_oracle_against_first_principles(n_fft, hop, win, N):
    """torch.stft/istft as the reference calls them == explicit float64 DFT / reflect / overlap-add."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal(N)
    cfg = dict(sampling_rate=N, n_fft=n_fft, hop_length=hop, win_length=win, audio_length=1)
    X, mag, ph = R.compute_stft(torch.from_numpy(x).float(), **cfg)
    Xd = dft64.stft(x.astype(np.float32), n_fft, hop, win)
    assert relerr(X.numpy(), Xd) < 2e-6
    y = R.compute_invert_stft(X, **cfg).numpy()
    yd = dft64.istft(Xd, n_fft, hop, win, length=N)
    assert relerr(y, yd) < 5e-6
    # the phase-preserving gain form used by the CUDA kernels == the reference's polar form
    mask = rng.random(Xd.shape)
    rel, irr = R.mask_apply_log1p(mag, ph, torch.from_numpy(mask).float())
    rd, idd = dft64.mask_apply(Xd, mask)
    assert relerr(rel.numpy(), rd) < 5e-6 and relerr(irr.numpy(), idd) < 5e-6


def test_istft_nola_and_errors():
    with pytest.raises(ValueError, match="waveform must be 1D"):
        R.compute_stft(torch.zeros(1, 2, 3))
    with pytest.raises(ValueError, match="ISTFT expects complex input"):
        R.compute_invert_stft(torch.zeros(3, 4))
    with pytest.raises(RuntimeError):
        dft64.istft(np.zeros((33, 4), complex), 64, 64, 16, length=100)  # holes between frames


def test_dft64_metric_scores_match_oracle():
    g = golden("lmac_metrics.npz")
    p, th, q = (g[f"rand_{k}"] for k in ("p", "theta", "q"))
    s = R.lmac_scores(*(torch.from_numpy(a) for a in (p, th, q))).numpy()
    d = dft64.lmac_scores(p, th, q)
    np.testing.assert_allclose(s, d, atol=2e-4)


def test_align_shift_against_reference_golden():
    """oracle align_shift vs the unmodified reference align_waveforms (hifigan.py:113-136, lifted by make_golden)."""
    g = golden("align.npz")
    for i in range(3):
        ref, deg = torch.from_numpy(g[f"ref{i}"]), torch.from_numpy(g[f"deg{i}"])
        shift = R.align_shift(ref, deg)
        assert shift == int(g[f"delay{i}"])
        ra = g[f"ref_aligned{i}"]
        assert ra.shape[:2] == (1, 1)
        want = ref[shift:shift + ra.shape[-1]] if shift > 0 else ref[:ra.shape[-1]]
        assert np.array_equal(ra[0, 0], want.numpy())
