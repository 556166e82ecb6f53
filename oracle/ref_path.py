"""torch-CPU restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).

Every function names the reference lines it follows (paths relative to
/root/reference).  The reference's arithmetic on this path is a sequence of
torch library calls, so the restatement makes the same calls with the same
arguments on CPU tensors; nothing here is tuned or re-derived.

Pinned by tests/test_oracle_golden.py against
  * tests/golden/*.npz - outputs of the unmodified reference modules executed
    in the build container (oracle/make_golden.py), and
  * oracle/dft64.py     - first-principles float64 definitions.
"""
from __future__ import annotations

import math
import warnings

import torch
import torch.nn.functional as F

EPS = 1e-10  # LMAC_metrics.py:28


# --------------------------------------------------------------------------
# audioprocessor.py
# --------------------------------------------------------------------------
def fit_length(wave: torch.Tensor, length: int) -> torch.Tensor:
    """Right-pad with zeros or crop the last dim to ``length`` samples.

    audioprocessor.py:83-98 (same branch for 1-D and 2-D input) and :56-62.
    """
    if wave.dim() not in (1, 2):
        raise ValueError("waveform must be 1D (single) or 2D (batched waveforms)")  # :100
    cur = wave.shape[-1]
    if cur < length:
        return F.pad(wave, (0, length - cur))
    return wave[..., :length]


def compute_stft(wave, *, sampling_rate=16000, n_fft=1024, hop_length=322, win_length=644,
                 audio_length=5, window=None):
    """audioprocessor.py:82-112 -> (X complex64 [.,F,T], |X|, angle(X)).

    ``window=None`` is the reference call (rectangular ``win_length`` window,
    centre-padded to ``n_fft`` by torch); a window tensor reproduces the
    hifigan.py:189-204 call sites.
    """
    wave = fit_length(wave, int(audio_length * sampling_rate))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        X = torch.stft(wave, n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                       window=window, return_complex=True)
    return X, X.abs(), X.angle()


def compute_invert_stft(spec, *, sampling_rate=16000, n_fft=1024, hop_length=322,
                        win_length=644, audio_length=5, window=None, use_length=True):
    """audioprocessor.py:117-131; ``use_length=False`` is hifigan.py:223-225."""
    if not torch.is_complex(spec):
        raise ValueError("ISTFT expects complex input!")  # :119
    length = int(audio_length * sampling_rate) if use_length else None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return torch.istft(spec, n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                           window=window, length=length)


def zero_mean_unit_var_norm(x):
    """classifier_embedder.py:59-63 (unbiased std, eps added to the std)."""
    mu = x.mean(dim=-1, keepdim=True)
    sd = x.std(dim=-1, keepdim=True)
    return (x - mu) / (sd + 1e-7)


def mel_transform(wave, *, sampling_rate=16000, n_fft=1024, hop_length=322, win_length=644,
                  n_mels=80):
    """audioprocessor.py:38-44: torchaudio MelSpectrogram defaults
    (hann(win_length) window, power 2, htk scale, no filterbank norm, centre/reflect)."""
    import torchaudio.transforms as T
    tr = T.MelSpectrogram(sample_rate=sampling_rate, n_fft=n_fft, hop_length=hop_length,
                          win_length=win_length, n_mels=n_mels)
    return tr(wave)


# --------------------------------------------------------------------------
# mask arithmetic
# --------------------------------------------------------------------------
def extend_mask(mask, F_bins, T_frames):
    """Zero-extend a [B,F',T'] mask to [B,F,T] (our documented convention for
    F' <= F, T' <= T; the reference itself only runs with full-size masks,
    SURVEY.md section 2.3 item 3)."""
    B, Fm, Tm = mask.shape
    if Fm == F_bins and Tm == T_frames:
        return mask
    out = mask.new_zeros(B, F_bins, T_frames)
    out[:, :Fm, :Tm] = mask
    return out


def mask_apply_log1p(mag, phase, mask):
    """LMAC_metrics.py:136-143 and :151-153 (== streamlit_controlled_study.py:173-183).

    rel = expm1(mask * log1p(mag)) * exp(1j*phase); irr likewise with (1-mask).
    """
    lm = torch.log1p(mag)
    rel = torch.expm1(mask * lm) * torch.exp(1j * phase)
    irr = torch.expm1((1 - mask) * lm) * torch.exp(1j * phase)
    return rel, irr


def mask_apply_linear(mag, phase, mask):
    """loss_function.py:36-45: (mask * mag) * exp(1j*phase), ((1-mask) * mag) * exp(1j*phase)."""
    ph = torch.exp(1j * phase)
    return (mask * mag) * ph, ((1 - mask) * mag) * ph


def mask_head(y1, weight, bias):
    """addvisor.py:57-60,82: Conv2d(32,1,1) + Sigmoid on [B,32,F',T'] -> [B,1,F',T']."""
    return torch.sigmoid(F.conv2d(y1, weight, bias))


def explain(wave, mask, *, mode="log1p", normalize=False, outside="drop", **cfg):
    """The LMAC_metrics.py:125-158 loop body without the classifier:
    STFT -> mask-apply (mask and 1-mask) -> 2 x iSTFT [-> normaliser].

    A mask [B,F',T'] smaller than the grid: ``outside="drop"`` follows the reference's crop of magnitude and
    phase to the mask's extent (LMAC_metrics.py:136-139, loss_function.py:36-41) - both masked spectra exist
    on [F',T'] only and are zero-padded back to [F,T] so that torch.istft accepts them; ``"keep_irr"``
    zero-extends the mask instead.

    Returns (rel_wave, irr_wave)."""
    _, mag, phase = compute_stft(wave, **cfg)
    fn = mask_apply_log1p if mode == "log1p" else mask_apply_linear
    Fm, Tm = mask.shape[-2], mask.shape[-1]
    if outside == "drop" and (Fm, Tm) != tuple(mag.shape[-2:]):
        r, i = fn(mag[:, :Fm, :Tm], phase[:, :Fm, :Tm], mask)
        rel = torch.zeros(mag.shape, dtype=r.dtype)
        irr = torch.zeros(mag.shape, dtype=i.dtype)
        rel[:, :Fm, :Tm], irr[:, :Fm, :Tm] = r, i
    else:
        m = extend_mask(mask, mag.shape[-2], mag.shape[-1])
        rel, irr = fn(mag, phase, m)
    icfg = {k: v for k, v in cfg.items()}
    rel_w = compute_invert_stft(rel, **icfg)
    irr_w = compute_invert_stft(irr, **icfg)
    if normalize:
        rel_w, irr_w = zero_mean_unit_var_norm(rel_w), zero_mean_unit_var_norm(irr_w)
    return rel_w, irr_w


# --------------------------------------------------------------------------
# classifier head + LMAC metrics
# --------------------------------------------------------------------------
def logreg(x, coef, intercept):
    """classifier_embedder.py:21-38: Linear(1920,1) -> (logits, sigmoid(logits))."""
    logits = F.linear(x, coef, intercept)
    return logits, torch.sigmoid(logits)


def score_for_predicted_class(p):
    """LMAC_metrics.py:43-45."""
    pred = (p > 0.5).float()
    return pred * p + (1 - pred) * (1 - p)


def fidelity(theta_out, predictions, threshold=0.5):
    """LMAC_metrics.py:31-38 -> [N,1] float."""
    return ((predictions > threshold).long() == (theta_out > threshold).long()).float()


def faithfulness(predictions, predictions_masked):
    """LMAC_metrics.py:48-52 -> [N]."""
    return ((predictions - predictions_masked) * torch.sign(predictions - 0.5)).squeeze(dim=1)


def average_drop(theta_out, predictions):
    """LMAC_metrics.py:55-59 -> [N]."""
    pc = score_for_predicted_class(predictions.squeeze(1))
    oc = score_for_predicted_class(theta_out.squeeze(1))
    return (F.relu(pc - oc) / (pc + EPS)) * 100


def average_increase(theta_out, predictions):
    """LMAC_metrics.py:62-66 -> [N]."""
    pc = score_for_predicted_class(predictions.squeeze(1))
    oc = score_for_predicted_class(theta_out.squeeze(1))
    return (oc > pc).float() * 100


def average_gain(theta_out, predictions):
    """LMAC_metrics.py:69-73 -> [N]."""
    pc = score_for_predicted_class(predictions.squeeze(1))
    oc = score_for_predicted_class(theta_out.squeeze(1))
    return (F.relu(oc - pc) / (1 - pc + EPS)) * 100


def lmac_scores(predictions, theta_out, masked_predictions):
    """Per-sample [N,5] = (FF, Fid, AD, AI, AG) from three [N,1] probability tensors."""
    return torch.stack([
        faithfulness(predictions, masked_predictions),
        fidelity(theta_out, predictions).squeeze(1),
        average_drop(theta_out, predictions),
        average_increase(theta_out, predictions),
        average_gain(theta_out, predictions),
    ], dim=1)


def lmac_sums(predictions, theta_out, masked_predictions):
    """float64 sums of the five per-sample scores plus the sample count
    (the six partials the multi-GPU path all-reduces)."""
    s = lmac_scores(predictions, theta_out, masked_predictions).double().sum(dim=0)
    return torch.cat([s, torch.tensor([float(predictions.shape[0])], dtype=torch.float64)])


def lmac_means(predictions, theta_out, masked_predictions):
    """LMAC_metrics.py:160-172: the five printed means, in print order
    (faithfulness, fidelity, average drop, average increase, average gain)."""
    return lmac_scores(predictions, theta_out, masked_predictions).mean(dim=0)


# --------------------------------------------------------------------------
# captum path (time-domain mask) and band swapping
# --------------------------------------------------------------------------
def td_mask(wave, attribution):
    """captum_saliency.py:136-143: m = |attr| / (max|attr| + 1e-8) over the whole
    tensor passed in (one clip there); returns (mask, wave*m, wave*(1-m))."""
    a = attribution.abs()
    m = a / (a.max() + 1e-8)
    return m, wave * m, wave * (1 - m)


def band_swap(spec_real, spec_voc, start_hz, end_hz, f_top=8000.0):
    """train_logReg_swapping.py:64-75 / hifigan.py:206-214: copy the rows whose
    centre frequency lies in [start, end) from the vocoded STFT into the real one.
    ``freqs = linspace(0, f_top, F)``."""
    Fb = spec_real.shape[-2]
    freqs = torch.linspace(0, f_top, Fb)
    rows = (freqs >= start_hz) & (freqs < end_hz)
    out = spec_real.clone()
    out[..., rows, :] = spec_voc[..., rows, :]
    return out


def align_shift(ref_wav, deg_wav):
    """hifigan.py:113-136 cross-correlation arg-max shift (direct O(N^2) conv1d)."""
    ref = ref_wav.reshape(1, 1, -1)
    deg = deg_wav.reshape(1, 1, -1)
    pad = deg.shape[-1]
    cc = F.conv1d(F.pad(ref, (pad, pad)), deg)
    return int(torch.argmax(cc).item()) - pad


def n_frames(n_samples: int, hop: int) -> int:
    return 1 + n_samples // hop


def n_bins(n_fft: int) -> int:
    return n_fft // 2 + 1


__all__ = [n for n in dir() if not n.startswith("_") and n not in ("math", "warnings", "torch", "F")]
